#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_fwd.py --iters 2 --bwd --quant > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python tools/prof_fwd.py --iters 2 --bwd --quant > gpurun_out/ncu_launches.log 2>&1
echo "exit=$?"; cat gpurun_out/plain.log
