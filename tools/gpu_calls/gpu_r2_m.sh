#!/bin/bash
mkdir -p gpurun_out
python tools/exp_stagger.py > gpurun_out/r2m_stagger.jsonl 2> gpurun_out/r2m_stagger.err; cat gpurun_out/r2m_stagger.jsonl; tail -3 gpurun_out/r2m_stagger.err
