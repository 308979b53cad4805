#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -4 gpurun_out/r2b_pytest.log
bash tools/run_variants.sh
ncu --set full --clock-control none --import-source on -k regex:k_hover_step -s 60 -c 1 -o gpurun_out/r2b_hover python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_ncu_full.log 2>&1
python bench.py --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -c 300 gpurun_out/r2b_bench.json
