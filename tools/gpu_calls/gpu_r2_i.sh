#!/bin/bash
# round 2, call I: new parity tests, the default bench line, the reference arm, the other workloads, launch list + ncu captures
mkdir -p gpurun_out
T=r2i
python -m pytest tests/test_timed_path_parity.py tests/test_dogfight_split.py tests/test_rocket.py -m gpu -q -s > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
grep -E "timed-path|passed|failed|rc=|FAILED|Error" gpurun_out/${T}_pytest.log | tail -30
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -3 gpurun_out/${T}_bench.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err
python tools/bench_workloads.py --steps 100 > gpurun_out/${T}_workloads.jsonl 2>> gpurun_out/${T}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 20 --warmup 3 --repeats 1 --no-cpu-baseline > gpurun_out/${T}_ncu_bench.log 2>&1
ncu --set full --clock-control none -k regex:k_hover_step -s 60 -c 1 -o gpurun_out/${T}_hover python bench.py --steps 20 --warmup 3 --repeats 1 --no-cpu-baseline > gpurun_out/${T}_ncu_full.log 2>&1
ncu --set full --clock-control none -k regex:"k_(fwwp|land|df|qxwp)_step" -c 12 -o gpurun_out/${T}_other python tools/bench_workloads.py --steps 2 --warmup 1 > gpurun_out/${T}_ncu_other.log 2>&1
PYFLYT_B200_LIB=$PWD/pyflyt_b200/lib/variants/fwl4/libpyflyt_b200.so ncu --set full --clock-control none --import-source on -k regex:k_fwwp_step -s 24 -c 3 -o gpurun_out/${T}_fwl4 python tools/bench_workloads.py --only fixedwing-waypoints --steps 5 --warmup 12 > gpurun_out/${T}_ncu_fwl4.log 2>&1
ls -la gpurun_out | grep ${T}; du -sh gpurun_out
