#!/bin/bash
# batch-norm kernel iteration: parity tests, then the ResNet-20 step at two batch sizes
mkdir -p gpurun_out
python -m pytest tests/test_gpu_bn.py tests/test_gpu_resnet20.py -m gpu -q -x 2>&1 | tail -5
python tools/train_bench.py --batch 256 2>&1 | tail -1
python tools/train_bench.py --batch 2048 --steps 10 2>&1 | tail -1
