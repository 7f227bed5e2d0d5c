#!/bin/bash
timeout 300 python tools/time_bwd.py --only v2 2>&1 | grep "avg\|grad_w"
timeout 900 python -m pytest tests/test_gpu_v2.py tests/test_gpu_matrix.py tests/test_gpu_parity.py tests/test_gpu_resnet20.py -m gpu -q -x 2>&1 | tail -2
for b in 256 2048; do timeout 600 python tools/train_bench.py --batch $b --steps 10 2>/dev/null | tail -1 | cut -c1-180; done
