#!/bin/bash
mkdir -p gpurun_out
T=r2q
python tools/exp_rotation.py > gpurun_out/${T}_rotation_pdl.jsonl 2> gpurun_out/${T}_rotation.err; cat gpurun_out/${T}_rotation_pdl.jsonl; tail -3 gpurun_out/${T}_rotation.err
PFB_PDL=0 python tools/exp_rotation.py > gpurun_out/${T}_rotation_nopdl.jsonl 2>> gpurun_out/${T}_rotation.err; cat gpurun_out/${T}_rotation_nopdl.jsonl
python -m pytest tests/test_gpu_parity.py tests/test_timed_path_parity.py tests/test_ma_quadx_hover.py -m gpu -q -s > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
grep -E "timed-path|passed|failed|rc=|FAILED|Error" gpurun_out/${T}_pytest.log | tail -30 | cut -c1-400
python bench.py --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -5 gpurun_out/${T}_bench.err; python -c "
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
c=d['config']
print('value',d['value'],'ms/step', d['ms_per_step'], 'blocks', c['block_ms'], 'flushed', c['ms_per_step_l2_flushed_event_pairs'], 'warm', c['ms_per_step_l2_warm'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['kernel_avg_us'], 'launches', d['gpu_launches'])
"
