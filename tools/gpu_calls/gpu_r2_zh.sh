#!/bin/bash
# A/B: Fixedwing-Waypoints step with the surfaces through ONE rolled copy (instruction-cache footprint) vs the one-basic-block FULL path;
# ncu of the rolled variant; ncu summaries of the rocket / dogfight step kernels (missing from the final call)
mkdir -p gpurun_out /tmp/ncu
T=r2zh
: > gpurun_out/${T}_fw_ab.jsonl
for rep in 1 2; do
  for lib in default fwroll; do
    if [ $lib = default ]; then unset PYFLYT_B200_LIB; else export PYFLYT_B200_LIB=$PWD/pyflyt_b200/lib/variants/$lib/libpyflyt_b200.so; fi
    echo "{\"lib\": \"$lib\"}" >> gpurun_out/${T}_fw_ab.jsonl
    python tools/bench_workloads.py --only fixedwing-waypoints --steps 100 >> gpurun_out/${T}_fw_ab.jsonl 2>> gpurun_out/${T}_fw_ab.err
  done
done
cut -c1-60 gpurun_out/${T}_fw_ab.jsonl; grep -o '"ms_per_step": [0-9.e-]*' gpurun_out/${T}_fw_ab.jsonl
export PYFLYT_B200_LIB=$PWD/pyflyt_b200/lib/variants/fwroll/libpyflyt_b200.so
ncu --set full --clock-control none -k regex:k_fwwp_step -s 24 -c 4 -o /tmp/ncu/fwroll python tools/bench_workloads.py --only fixedwing-waypoints --steps 5 --warmup 12 > gpurun_out/${T}_ncu_fwroll.log 2>&1
: > gpurun_out/${T}_fwroll_ncu_summary.txt
for w in 0 1 2 3; do python tools/ncu_summary.py /tmp/ncu/fwroll.ncu-rep $w >> gpurun_out/${T}_fwroll_ncu_summary.txt 2>/dev/null; done
unset PYFLYT_B200_LIB
ncu --set full --clock-control none -k regex:"k_(land|df)_step" -c 8 -o /tmp/ncu/landdf python tools/bench_workloads.py --steps 2 --warmup 1 > gpurun_out/${T}_ncu_landdf.log 2>&1
: > gpurun_out/${T}_land_df_ncu_summary.txt
for w in 0 1 2 3 4 5 6 7; do python tools/ncu_summary.py /tmp/ncu/landdf.ncu-rep $w >> gpurun_out/${T}_land_df_ncu_summary.txt 2>/dev/null; done
grep -E "^kernel|gpu__time_duration|grid_size|issue_active|no_instruction" gpurun_out/${T}_fwroll_ncu_summary.txt | cut -c1-120
