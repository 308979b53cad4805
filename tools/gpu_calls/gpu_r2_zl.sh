#!/bin/bash
# A/B: Fixedwing-Waypoints step with 4 / 8 warps per CTA (warps share an SM's instruction caches; fewer SMs busy) vs one warp per CTA
mkdir -p gpurun_out
T=r2zl
: > gpurun_out/${T}_fw_block_ab.jsonl
for rep in 1 2; do
  for lib in default b128 b256; do
    if [ $lib = default ]; then unset PYFLYT_B200_LIB; else export PYFLYT_B200_LIB=$PWD/pyflyt_b200/lib/variants/$lib/libpyflyt_b200.so; fi
    echo "{\"lib\": \"$lib\"}" >> gpurun_out/${T}_fw_block_ab.jsonl
    python tools/bench_workloads.py --only fixedwing-waypoints --steps 100 >> gpurun_out/${T}_fw_block_ab.jsonl 2>> gpurun_out/${T}_fw_block_ab.err
  done
done
grep -o '"lib": "[a-z0-9]*"\|"ms_per_step": [0-9.e-]*' gpurun_out/${T}_fw_block_ab.jsonl | paste - - ; tail -2 gpurun_out/${T}_fw_block_ab.err
