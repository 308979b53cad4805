#!/bin/bash
# round 2, call H: full GPU suite, the default bench line, the reference arm, the other workloads, launch list + ncu captures
mkdir -p gpurun_out
T=r2h
python tools/dbg_rocket_spare.py > gpurun_out/${T}_dbg.log 2>&1; tail -4 gpurun_out/${T}_dbg.log
python -m pytest tests -m gpu -q -s > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
grep -E "timed-path|north-star|passed|failed|rc=|FAILED" gpurun_out/${T}_pytest.log | tail -30
python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -3 gpurun_out/${T}_bench.err; python -c "
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print('value',d['value'],'ms/step', d['ms_per_step'], 'warm', d['config']['ms_per_step_l2_warm'],'e2e', d['e2e'], 'kernel', d['roofline']['kernel_avg_us'], d['roofline']['frac'], d['gpu_launches'], d.get('cpu_baseline',{}).get('value'))
"
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err
python tools/bench_workloads.py --steps 100 > gpurun_out/${T}_workloads.jsonl 2>> gpurun_out/${T}_bench.err; cut -c1-260 gpurun_out/${T}_workloads.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 20 --warmup 3 --repeats 1 --no-cpu-baseline > gpurun_out/${T}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_hover_step -s 60 -c 1 -o gpurun_out/${T}_hover python bench.py --steps 20 --warmup 3 --repeats 1 --no-cpu-baseline > gpurun_out/${T}_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_(fwwp|land|df|qxwp)_step" -c 12 -o gpurun_out/${T}_other python tools/bench_workloads.py --steps 2 --warmup 1 > gpurun_out/${T}_ncu_other.log 2>&1
ls -la gpurun_out | grep ${T}
