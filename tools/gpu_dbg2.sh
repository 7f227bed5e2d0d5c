#!/bin/bash
timeout 300 python tools/time_bwd.py --only v2 2>&1 | grep "dgrad\|grad_x"
timeout 900 python -m pytest tests/test_gpu_v2.py tests/test_gpu_matrix.py -m gpu -q -x 2>&1 | tail -2
