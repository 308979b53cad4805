#!/bin/bash
mkdir -p gpurun_out
python tools/dbg_rocket_env.py > gpurun_out/r2p_dbg_rocket_env.log 2>&1; tail -46 gpurun_out/r2p_dbg_rocket_env.log | cut -c1-600
python -m pytest tests/test_timed_path_parity.py -m gpu -q -s -k "dogfight" > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
grep -E "timed-path|passed|failed|rc=|FAILED|Error" gpurun_out/r2p_pytest.log | tail -10 | cut -c1-600
