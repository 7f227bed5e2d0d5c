#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/train_bench.py --batch 256 --steps 2 --warmup 3 --no-graph > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 1400 --csv --log-file gpurun_out/train_launches.csv \
    python tools/train_bench.py --batch 256 --steps 2 --warmup 3 --no-graph > gpurun_out/ncu_launches.log 2>&1
echo "exit=$?"; cat gpurun_out/plain.log | tail -2
