#!/bin/bash
# Whole GPU check: every test file in its own process under `timeout`, smoke, then the bench.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?"; tail -n 8 gpurun_out/$name.log; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
for f in tests/test_gpu_v2.py tests/test_gpu_parity.py tests/test_gpu_matrix.py tests/test_gpu_lsq_modules.py tests/test_gpu_bn.py tests/test_gpu_pipeline.py tests/test_gpu_resnet20.py tests/test_gpu_stochastic.py tests/test_launcher.py tests/test_cabi_and_host.py; do
  run r2_$(basename $f .py) python -m pytest $f -m gpu -q
done
run r2_smoke python __graft_entry__.py smoke
run r2_bench python bench.py
