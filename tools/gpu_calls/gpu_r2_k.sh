#!/bin/bash
mkdir -p gpurun_out
T=r2k
python -m pytest tests/test_gpu_parity.py tests/test_timed_path_parity.py tests/test_dogfight.py -m gpu -q -s -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
grep -E "timed-path|passed|failed|rc=|FAILED|Error" gpurun_out/${T}_pytest.log | tail -30
bash tools/run_variants.sh 2>&1 | tail -5
python bench.py --no-cpu-baseline --repeats 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -3 gpurun_out/${T}_bench.err; python -c "
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print('value',d['value'],'ms/step', d['ms_per_step'],'kernel', d['roofline']['kernel_avg_us'], d['roofline']['frac'], 'e2e', d['e2e']['value'])
"
