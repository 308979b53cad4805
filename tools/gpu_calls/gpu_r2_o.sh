#!/bin/bash
mkdir -p gpurun_out
T=r2o
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -5 gpurun_out/${T}_bench.err; python -c "
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
c=d['config']
print('value',d['value'],'ms/step', d['ms_per_step'], 'blocks', c['block_ms'], 'flushed', c['ms_per_step_l2_flushed_event_pairs'], 'warm', c['ms_per_step_l2_warm'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['kernel_avg_us'], 'launches', d['gpu_launches'], d['clocks'])
"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 20 --warmup 3 --repeats 1 --no-cpu-baseline > gpurun_out/${T}_ncu_bench.log 2>&1
python -m pytest tests -m gpu -q -s > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
grep -E "timed-path|passed|failed|rc=|FAILED|Error" gpurun_out/${T}_pytest.log | tail -30 | cut -c1-500
