#!/bin/bash
mkdir -p gpurun_out
T=r2l
PYFLYT_B200_LIB=$PWD/pyflyt_b200/lib/variants/tl/libpyflyt_b200.so python tools/exp_timeline.py > gpurun_out/${T}_timeline.jsonl 2> gpurun_out/${T}_timeline.err; cat gpurun_out/${T}_timeline.jsonl | cut -c1-700; tail -3 gpurun_out/${T}_timeline.err
python -m pytest tests/test_timed_path_parity.py -m gpu -q -s -k "rocket or dogfight" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
grep -E "timed-path|passed|failed|rc=|FAILED|Error" gpurun_out/${T}_pytest.log | tail -30 | cut -c1-600
