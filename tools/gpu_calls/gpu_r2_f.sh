#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
grep -E "timed-path|north-star|passed|failed|rc=|FAILED" gpurun_out/r2f_pytest.log | tail -30
python __graft_entry__.py smoke > gpurun_out/r2f_smoke.log 2>&1; tail -2 gpurun_out/r2f_smoke.log
python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -3 gpurun_out/r2f_bench.err; python -c "
import json
d=json.load(open('gpurun_out/r2f_bench.json'))
print('value',d['value'],'ms/step', d['ms_per_step'], 'warm', d['config']['ms_per_step_l2_warm'],'e2e', d['e2e']['value'], 'kernel', d['roofline']['kernel_avg_us'], d['roofline']['frac'], d['gpu_launches'], d.get('cpu_baseline',{}).get('value'))
"
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2f_bench_reference.json 2>> gpurun_out/r2f_bench.err
