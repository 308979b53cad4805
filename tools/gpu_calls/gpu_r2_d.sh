#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
grep -E "timed-path|north-star|passed|failed|rc=|Error|FAILED" gpurun_out/r2d_pytest.log | tail -30
python bench.py --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -3 gpurun_out/r2d_bench.err; python -c "
import json
d=json.load(open('gpurun_out/r2d_bench.json'))
print('value',d['value'],'ms/step', d['ms_per_step'],'blocks', d['config']['block_ms'], 'warm', d['config']['ms_per_step_l2_warm'],'e2e', d['e2e']['value'], 'kernel', d['roofline']['kernel_avg_us'], d['roofline']['frac'], d['gpu_launches'])
"
ncu --set full --clock-control none --import-source on -k regex:k_hover_step -s 60 -c 1 -o gpurun_out/r2d_hover python bench.py --steps 20 --warmup 3 --repeats 1 --no-cpu-baseline > gpurun_out/r2d_ncu_full.log 2>&1
