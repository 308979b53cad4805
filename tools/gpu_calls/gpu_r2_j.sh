#!/bin/bash
# round 2, call J: debug + experiments + the bench / profile artefacts (ncu reports are summarised ON THE BOX: they are too big to travel)
mkdir -p gpurun_out
T=r2j
python tools/dbg_rocket_crash.py > gpurun_out/${T}_dbg_crash.log 2>&1; tail -40 gpurun_out/${T}_dbg_crash.log | cut -c1-330
python tools/exp_mapped_waves.py > gpurun_out/${T}_mapped_waves.jsonl 2> gpurun_out/${T}_mapped_waves.err; cat gpurun_out/${T}_mapped_waves.jsonl; tail -3 gpurun_out/${T}_mapped_waves.err
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "masked_reset" > gpurun_out/${T}_pytest_masked.log 2>&1; tail -5 gpurun_out/${T}_pytest_masked.log
python -m pytest tests/test_timed_path_parity.py -m gpu -q -s -k "quadx_waypoints" > gpurun_out/${T}_pytest_qxwp.log 2>&1; grep -E "timed-path|passed|failed" gpurun_out/${T}_pytest_qxwp.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -3 gpurun_out/${T}_bench.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err
python tools/bench_workloads.py --steps 100 > gpurun_out/${T}_workloads.jsonl 2>> gpurun_out/${T}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 20 --warmup 3 --repeats 1 --no-cpu-baseline > gpurun_out/${T}_ncu_bench.log 2>&1
mkdir -p /tmp/ncu
ncu --set full --clock-control none -k regex:k_hover_step -s 60 -c 1 -o /tmp/ncu/hover python bench.py --steps 20 --warmup 3 --repeats 1 --no-cpu-baseline > gpurun_out/${T}_ncu_full.log 2>&1
python tools/ncu_summary.py /tmp/ncu/hover.ncu-rep > gpurun_out/${T}_k_hover_step_ncu_summary.txt 2>&1
ncu --set full --clock-control none -k regex:"k_(fwwp|land|df|qxwp)_step" -c 12 -o /tmp/ncu/other python tools/bench_workloads.py --steps 2 --warmup 1 > gpurun_out/${T}_ncu_other.log 2>&1
: > gpurun_out/${T}_other_step_kernels_ncu_summary.txt
for w in 0 1 2 3 4 5 6 7 8 9 10 11; do python tools/ncu_summary.py /tmp/ncu/other.ncu-rep $w >> gpurun_out/${T}_other_step_kernels_ncu_summary.txt 2>/dev/null; done
PYFLYT_B200_LIB=$PWD/pyflyt_b200/lib/variants/fwl4/libpyflyt_b200.so ncu --set full --clock-control none --import-source on -k regex:k_fwwp_step -s 24 -c 3 -o /tmp/ncu/fwl4 python tools/bench_workloads.py --only fixedwing-waypoints --steps 5 --warmup 12 > gpurun_out/${T}_ncu_fwl4.log 2>&1
for w in 0 1 2; do python tools/ncu_summary.py /tmp/ncu/fwl4.ncu-rep $w >> gpurun_out/${T}_fwl4_ncu_summary.txt 2>/dev/null; done
ncu -i /tmp/ncu/fwl4.ncu-rep --page source --csv --print-source sass > gpurun_out/${T}_fwl4_source.csv 2>/dev/null; ls -la gpurun_out/${T}_fwl4_source.csv
ls -la gpurun_out | grep ${T}; du -sh gpurun_out
