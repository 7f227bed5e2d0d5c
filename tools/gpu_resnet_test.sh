#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_resnet20.py -m gpu -q -s > gpurun_out/t_resnet.log 2>&1; echo "exit=$?"; tail -n 12 gpurun_out/t_resnet.log | cut -c1-600

timeout 600 python -m pytest tests -m gpu -q -x -k "cabi or module_matches" > gpurun_out/t_cabi.log 2>&1; echo "exit=$?"; tail -n 2 gpurun_out/t_cabi.log
