#!/bin/bash
mkdir -p gpurun_out
python tools/exp_rotation.py > gpurun_out/r2n_rotation.jsonl 2> gpurun_out/r2n_rotation.err; cat gpurun_out/r2n_rotation.jsonl; tail -3 gpurun_out/r2n_rotation.err
