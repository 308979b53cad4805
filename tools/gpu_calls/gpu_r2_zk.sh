#!/bin/bash
# source-level stall view of the one-basic-block Fixedwing-Waypoints step (where do the no_instruction stalls sit?)
mkdir -p gpurun_out /tmp/ncu
T=r2zk
ncu --set full --clock-control none --import-source on -k regex:k_fwwp_step -s 25 -c 1 -o /tmp/ncu/fw python tools/bench_workloads.py --only fixedwing-waypoints --steps 5 --warmup 12 > gpurun_out/${T}_ncu_fw.log 2>&1
python tools/ncu_summary.py /tmp/ncu/fw.ncu-rep > gpurun_out/${T}_k_fwwp_step_ncu_summary.txt 2>&1
ncu -i /tmp/ncu/fw.ncu-rep --page source --csv --print-source sass > gpurun_out/${T}_fw_source.csv 2>/dev/null
head -6 gpurun_out/${T}_k_fwwp_step_ncu_summary.txt; ls -la gpurun_out/${T}_fw_source.csv
