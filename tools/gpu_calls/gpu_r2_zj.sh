#!/bin/bash
# A/B on the FUSED rollout: substep with 2 basic blocks instead of 5 (contact test first, velocity clamp as selects, wind compiled out)
mkdir -p gpurun_out
T=r2zj
: > gpurun_out/${T}_fused_ab.jsonl
for rep in 1 2; do
  for lib in default e123; do
    if [ $lib = default ]; then unset PYFLYT_B200_LIB; else export PYFLYT_B200_LIB=$PWD/pyflyt_b200/lib/variants/$lib/libpyflyt_b200.so; fi
    PFB_FUSED_M=12 PFB_FUSED_CHUNKS=1,16 python tools/exp_fused.py >> gpurun_out/${T}_fused_ab.jsonl 2>> gpurun_out/${T}_fused_ab.err
  done
done
cut -c1-170 gpurun_out/${T}_fused_ab.jsonl; tail -2 gpurun_out/${T}_fused_ab.err
