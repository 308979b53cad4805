#!/bin/bash
# times the default build and every build under pyflyt_b200/lib/variants/ (one process each); results -> gpurun_out/variants.jsonl
mkdir -p gpurun_out
: > gpurun_out/variants.jsonl
python tools/time_variants.py >> gpurun_out/variants.jsonl 2>> gpurun_out/variants.err
for d in pyflyt_b200/lib/variants/*/; do
  PYFLYT_B200_LIB=$PWD/${d}libpyflyt_b200.so python tools/time_variants.py >> gpurun_out/variants.jsonl 2>> gpurun_out/variants.err
done
python - <<'PY'
import json
for l in open('gpurun_out/variants.jsonl'):
    d = json.loads(l)
    name = d['lib'].split('/')[-2] if '/' in d['lib'] else d['lib']
    print(f"{name:14s} cold(side) p50 {d['cold_side']['p50']:5.1f} cold(same) p50 {d['cold_same']['p50']:5.1f} warm(same) p50 {d['warm_same']['p50']:5.1f} loop {d['warm_same']['loop_us_per_step']:5.1f} | warm(side) loop {d['warm_side']['loop_us_per_step']:5.1f} | parity {d['parity']}")
PY
