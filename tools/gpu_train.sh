#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/train_bench.py --batch 256 --steps 10 --warmup 3 > gpurun_out/train.json 2> gpurun_out/train.err; echo "exit=$?"; cat gpurun_out/train.json; tail -n 8 gpurun_out/train.err
timeout 600 python tools/train_bench.py --batch 256 --steps 10 --warmup 3 --no-graph > gpurun_out/train_eager.json 2>> gpurun_out/train.err; cat gpurun_out/train_eager.json
