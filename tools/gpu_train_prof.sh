#!/bin/bash
# ncu launch list of ResNet-20 training steps (no CUDA graph, so that every kernel is a launch): BATCH env (default 256)
B=${BATCH:-256}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-1200} -c 4000 --csv --log-file gpurun_out/r02_launches_train_b$B.csv python tools/train_bench.py --batch $B --steps 2 --warmup 3 --no-graph > gpurun_out/ncu_train.log 2>&1
echo "ncu exit=$?"; wc -l gpurun_out/r02_launches_train_b$B.csv
