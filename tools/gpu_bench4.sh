#!/bin/bash
# parity subset + our bench arm (development helper)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "cabi or random_layer or full_size or module_matches or deterministic" > gpurun_out/t_sub.log 2>&1; echo "tests exit=$?"; tail -n 2 gpurun_out/t_sub.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
