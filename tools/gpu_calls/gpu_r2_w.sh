#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
ncu --set full --clock-control none -k regex:k_hover_rollout -s 3 -c 1 -o /tmp/ncu/rollout python tools/prof_fused.py > gpurun_out/r2w_ncu.log 2>&1
python tools/ncu_summary.py /tmp/ncu/rollout.ncu-rep > gpurun_out/r2w_k_hover_rollout_ncu_summary.txt 2>&1; cat gpurun_out/r2w_k_hover_rollout_ncu_summary.txt
