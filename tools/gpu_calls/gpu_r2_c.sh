#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
grep -E "timed-path|passed|failed|rc=|Error|assert" gpurun_out/r2c_pytest.log | tail -12
bash tools/run_variants.sh
python bench.py --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; python -c "
import json
d=json.load(open('gpurun_out/r2c_bench.json'))
print('value',d['value'],'ms/step', d['ms_per_step'],'warm', d['config']['ms_per_step_l2_warm'],'inline', d['config']['ms_per_step_inline_resets'],'e2e', d['e2e']['value'], 'kernel', d['roofline']['kernel_avg_us'], d['gpu_launches'])
"
ncu --set full --clock-control none --import-source on -k regex:k_hover_step -s 60 -c 1 -o gpurun_out/r2c_hover python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_ncu_full.log 2>&1
