#!/bin/bash
# round 2, GPU call A: correctness of the restructured hover kernel + first measurements
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 600 gpurun_out/r2a_bench.json
python tools/exp_cold.py > gpurun_out/r2a_cold.log 2>&1; cat gpurun_out/r2a_cold.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_hover_step -s 60 -c 2 -o gpurun_out/r2a_hover python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_ncu_full.log 2>&1
ls -la gpurun_out | tail -8
