#!/bin/bash
mkdir -p gpurun_out
T=r2t
python -m pytest tests/test_gpu_parity.py tests/test_timed_path_parity.py -m gpu -q -s -k "fused" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; grep -E "fused|passed|failed|rc=|Error|assert" gpurun_out/${T}_pytest.log | tail -14 | cut -c1-400
python tools/exp_fused.py > gpurun_out/${T}_fused.jsonl 2> gpurun_out/${T}_fused.err; cat gpurun_out/${T}_fused.jsonl; tail -3 gpurun_out/${T}_fused.err
