#!/bin/bash
# 2 GPUs: the bench line under torchrun (as the driver launches it) + the 2-rank split-dogfight tests
mkdir -p gpurun_out
T=r2r
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err; tail -5 gpurun_out/${T}_bench_n2.err
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench_n2.json'))
c=d['config']
print('value',d['value'],'ms/step', d['ms_per_step'], 'n_gpus', d['n_gpus'], 'strong', c.get('value_strong_65536_total'), 'split', json.dumps(c.get('dogfight_split'))[:600], 'e2e', d['e2e']['value'])
"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > gpurun_out/${T}_bench_ref_n2.json 2>> gpurun_out/${T}_bench_n2.err; cut -c1-300 gpurun_out/${T}_bench_ref_n2.json
python -m pytest tests/test_dogfight_split.py -m gpu -q > gpurun_out/${T}_pytest_split.log 2>&1; tail -4 gpurun_out/${T}_pytest_split.log
