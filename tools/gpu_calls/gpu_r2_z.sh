#!/bin/bash
# A/B of the one-basic-block surface code (FULL) for the aerodynamic-surface step kernels + the GPU suite on the default build
mkdir -p gpurun_out
: > gpurun_out/r2z_workloads.jsonl
for lib in default nofull full_r8; do
  if [ $lib = default ]; then unset PYFLYT_B200_LIB; else export PYFLYT_B200_LIB=$PWD/pyflyt_b200/lib/variants/$lib/libpyflyt_b200.so; fi
  echo "{\"lib\": \"$lib\"}" >> gpurun_out/r2z_workloads.jsonl
  python tools/bench_workloads.py --steps 100 --warmup 5 >> gpurun_out/r2z_workloads.jsonl 2>> gpurun_out/r2z_workloads.err
done
unset PYFLYT_B200_LIB
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2z_pytest.log
python - <<'PY'
import json
for l in open('gpurun_out/r2z_workloads.jsonl'):
    d = json.loads(l)
    print(d.get('lib') or (d['workload'][:40], round(d['ms_per_step']*1e3, 2)))
PY
