#!/bin/bash
# spare records staged through shared memory (cp.async): GPU suite + bench
mkdir -p gpurun_out
T=r2zc
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -3 gpurun_out/${T}_bench.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2zc_bench.json').read().strip().splitlines()[-1])
c=l['config']
print('value',l['value'],'us',l['ms_per_step']*1e3,'flushed',c['ms_per_step_l2_flushed_event_pairs']*1e3,'warm',c['ms_per_step_l2_warm']*1e3,'e2e',l['e2e']['value'],'frac',l['roofline']['frac'])
print('fused',c['rollout_fused']['us_per_step'],c['rollout_fused']['frac_hbm_roofline'])
PY
