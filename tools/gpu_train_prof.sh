#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 4000 --csv --log-file gpurun_out/r02_launches_train_b256.csv python tools/train_bench.py --batch 256 --steps 2 --warmup 3 --no-graph > gpurun_out/ncu_train.log 2>&1
echo "ncu exit=$?"; wc -l gpurun_out/r02_launches_train_b256.csv
