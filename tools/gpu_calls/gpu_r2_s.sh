#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
echo "new"; python tools/exp_rotation.py 2>/dev/null | grep -E '"batches": (1|12|16),' | cut -c1-120
echo "base"; PYFLYT_B200_LIB=$PWD/pyflyt_b200/lib/variants/base/libpyflyt_b200.so python tools/exp_rotation.py 2>/dev/null | grep -E '"batches": (1|12|16),' | cut -c1-120
done
