#!/bin/bash
# round 2, final call: GPU suite, bench (both arms), workloads, ncu launch list + full captures of the final kernels (summarised on the box)
mkdir -p gpurun_out /tmp/ncu
T=r2fin
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench.err
python bench.py > gpurun_out/${T}_bench.json 2>> gpurun_out/${T}_bench.err; tail -3 gpurun_out/${T}_bench.err
python tools/bench_workloads.py --steps 100 > gpurun_out/${T}_workloads.jsonl 2>> gpurun_out/${T}_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 20 --warmup 3 --repeats 1 --no-cpu-baseline > gpurun_out/${T}_ncu_bench.log 2>&1
ncu --set full --clock-control none -k regex:k_hover_step -s 60 -c 1 -o /tmp/ncu/hover python bench.py --steps 20 --warmup 3 --repeats 1 --no-cpu-baseline > gpurun_out/${T}_ncu_full.log 2>&1
python tools/ncu_summary.py /tmp/ncu/hover.ncu-rep > gpurun_out/${T}_k_hover_step_ncu_summary.txt 2>&1
ncu --set full --clock-control none -k regex:k_hover_rollout -s 3 -c 1 -o /tmp/ncu/rollout python tools/prof_fused.py > gpurun_out/${T}_ncu_rollout.log 2>&1
python tools/ncu_summary.py /tmp/ncu/rollout.ncu-rep > gpurun_out/${T}_k_hover_rollout_ncu_summary.txt 2>&1
ncu --set full --clock-control none -k regex:"k_(fwwp|land|df|qxwp)_step" -c 12 -o /tmp/ncu/other python tools/bench_workloads.py --steps 2 --warmup 1 > gpurun_out/${T}_ncu_other.log 2>&1
: > gpurun_out/${T}_other_step_kernels_ncu_summary.txt
for w in 0 1 2 3 4 5 6 7 8 9 10 11; do python tools/ncu_summary.py /tmp/ncu/other.ncu-rep $w >> gpurun_out/${T}_other_step_kernels_ncu_summary.txt 2>/dev/null; done
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2fin_bench.json').read().strip().splitlines()[-1])
c=l['config']
print('value',l['value'],'us',l['ms_per_step']*1e3,'flushed',c['ms_per_step_l2_flushed_event_pairs']*1e3,'warm',c['ms_per_step_l2_warm']*1e3,'e2e',l['e2e']['value'],'frac',l['roofline']['frac'],'cpu',l['cpu_baseline']['value'])
print('fused',c['rollout_fused']['us_per_step'],c['rollout_fused']['frac_hbm_roofline'])
PY
ls -la gpurun_out | grep ${T}
