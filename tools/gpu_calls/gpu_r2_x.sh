#!/bin/bash
mkdir -p gpurun_out
T=r2x
python -m pytest tests -m gpu -q -s > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
grep -E "passed|failed|rc=|FAILED|Error" gpurun_out/${T}_pytest.log | tail -12 | cut -c1-300
python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -5 gpurun_out/${T}_bench.err; python -c "
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
c=d['config']
print('value',d['value'],'ms/step', d['ms_per_step'], 'flushed', c['ms_per_step_l2_flushed_event_pairs'], 'warm', c['ms_per_step_l2_warm'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], 'launches', d['gpu_launches'])
print('fused', c['rollout_fused']['env_steps_per_s'], c['rollout_fused']['us_per_step'], c['rollout_fused']['frac_hbm_roofline'], c['rollout_fused']['launches_per_block'])
"
