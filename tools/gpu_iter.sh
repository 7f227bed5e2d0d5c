#!/bin/bash
# quick iteration: backward/forward parity subset + launch list
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?"; tail -n 4 gpurun_out/$name.log; }
run t_cabi   python -m pytest tests -m gpu -q -x -k "cabi or random_layer or full_size or lsq or module_matches"
timeout 300 python tools/prof_fwd.py --iters 3 --bwd --quant > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python tools/prof_fwd.py --iters 2 --bwd --quant > gpurun_out/ncu_launches.log 2>&1
echo "exit=$?"; cat gpurun_out/plain.log
