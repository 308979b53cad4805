#!/bin/bash
# A/B: done-list atomic issued right after the integration (inline PTX) vs at the end; timeline of the variant
mkdir -p gpurun_out
T=r2zf
: > gpurun_out/${T}_rotation.jsonl
for rep in 1 2; do
  for lib in default ea; do
    if [ $lib = default ]; then unset PYFLYT_B200_LIB; else export PYFLYT_B200_LIB=$PWD/pyflyt_b200/lib/variants/$lib/libpyflyt_b200.so; fi
    PFB_ROTATION_M=1,12 python tools/exp_rotation.py >> gpurun_out/${T}_rotation.jsonl 2>> gpurun_out/${T}_rotation.err
  done
done
cat gpurun_out/${T}_rotation.jsonl
PYFLYT_B200_LIB=$PWD/pyflyt_b200/lib/variants/ea_tl/libpyflyt_b200.so python tools/exp_timeline.py > gpurun_out/${T}_timeline.jsonl 2> gpurun_out/${T}_timeline.err; tail -2 gpurun_out/${T}_timeline.jsonl | cut -c1-1500
