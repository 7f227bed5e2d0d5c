#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_matrix.py tests/test_gpu_v2.py -m gpu -q -x 2>&1 | tail -3
