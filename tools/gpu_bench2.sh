#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "tests exit=$?"; tail -n 3 gpurun_out/t_all.log
