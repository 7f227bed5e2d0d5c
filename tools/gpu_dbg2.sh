#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_bn.py -m gpu -q -x -k "resnet20_step" 2>&1 | tail -12
