#!/bin/bash
# tensor-kernel iteration: layer parity tests, then layer times at the three ResNet-20 geometries and the microbench layer
python -m pytest tests/test_gpu_v2.py tests/test_gpu_matrix.py -m gpu -q -x 2>&1 | tail -2
for cfg in "16 32 1024" "32 16 1024" "64 8 1024" "64 32 256"; do
  set -- $cfg; echo -n "ch=$1 hw=$2 B=$3  "; python tools/prof_v2.py --channels $1 --hw $2 --batch $3 --iters 4 --time 2>&1 | grep forward
done
