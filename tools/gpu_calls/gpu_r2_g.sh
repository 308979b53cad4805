#!/bin/bash
mkdir -p gpurun_out
python tools/dbg_rocket_spare.py > gpurun_out/r2g_dbg.log 2>&1; cat gpurun_out/r2g_dbg.log | tail -6
python -m pytest tests/test_timed_path_parity.py tests/test_contact_response.py tests/test_wind.py -m gpu -q -s > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
grep -E "timed-path|passed|failed|rc=|FAILED|Error" gpurun_out/r2g_pytest.log | tail -12
python bench.py --no-cpu-baseline --repeats 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; tail -3 gpurun_out/r2g_bench.err; python -c "
import json
d=json.load(open('gpurun_out/r2g_bench.json'))
print('value',d['value'],'e2e', d['e2e'])
"
