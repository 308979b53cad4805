#!/bin/bash
# source-level stall sampling of the fused rollout loop and of the one-basic-block fixed-wing step; GPU suite on the new default build
mkdir -p gpurun_out /tmp/ncu
T=r2zb
ncu --set full --clock-control none --import-source on -k regex:k_hover_rollout -s 3 -c 1 -o /tmp/ncu/rollout python tools/prof_fused.py > gpurun_out/${T}_ncu_rollout.log 2>&1
python tools/ncu_summary.py /tmp/ncu/rollout.ncu-rep > gpurun_out/${T}_k_hover_rollout_ncu_summary.txt 2>&1
ncu -i /tmp/ncu/rollout.ncu-rep --page source --csv --print-source sass > gpurun_out/${T}_rollout_source.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:k_fwwp_step -s 24 -c 1 -o /tmp/ncu/fw python tools/bench_workloads.py --only fixedwing-waypoints --steps 5 --warmup 12 > gpurun_out/${T}_ncu_fw.log 2>&1
python tools/ncu_summary.py /tmp/ncu/fw.ncu-rep > gpurun_out/${T}_k_fwwp_step_ncu_summary.txt 2>&1
ncu -i /tmp/ncu/fw.ncu-rep --page source --csv --print-source sass > gpurun_out/${T}_fw_source.csv 2>/dev/null
ls -la gpurun_out | grep ${T}
python tools/bench_workloads.py --steps 100 --warmup 5 > gpurun_out/${T}_workloads.jsonl 2> gpurun_out/${T}_workloads.err
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
cut -c1-200 gpurun_out/${T}_workloads.jsonl
