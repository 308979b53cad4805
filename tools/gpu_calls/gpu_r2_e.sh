#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s -x > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
grep -E "timed-path|north-star|passed|failed|rc=|FAILED" gpurun_out/r2e_pytest.log | tail -20
: > gpurun_out/r2e_fw_lanes.jsonl
echo "lanes=4 (default)" >> gpurun_out/r2e_fw_lanes.jsonl
python tools/bench_workloads.py --only fixedwing-waypoints --steps 200 >> gpurun_out/r2e_fw_lanes.jsonl 2>> gpurun_out/r2e_fw.err
for v in fwl1 fwl2 fwl8; do
  echo "variant $v" >> gpurun_out/r2e_fw_lanes.jsonl
  PYFLYT_B200_LIB=$PWD/pyflyt_b200/lib/variants/$v/libpyflyt_b200.so python tools/bench_workloads.py --only fixedwing-waypoints --steps 200 >> gpurun_out/r2e_fw_lanes.jsonl 2>> gpurun_out/r2e_fw.err
done
python - <<'PY'
import json
for l in open('gpurun_out/r2e_fw_lanes.jsonl'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('   ', round(d['ms_per_step']*1e3,2), 'us/step', f"{d['value']:.3e}")
    else: print(l)
PY
ncu --set full --clock-control none --import-source on -k regex:k_fwwp_step -s 20 -c 1 -o gpurun_out/r2e_fwwp python tools/bench_workloads.py --only fixedwing-waypoints --steps 5 --warmup 12 > gpurun_out/r2e_ncu.log 2>&1
python tools/bench_workloads.py --steps 100 > gpurun_out/r2e_workloads.jsonl 2>> gpurun_out/r2e_fw.err
