#!/bin/bash
# eight-rank record: bench.py (driver's launch line) + ResNet-20 training at global batch 2048 (256 per GPU) and 2048 per GPU
mkdir -p gpurun_out
N=${N:-8}
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $L bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N exit=$?"; cut -c1-400 gpurun_out/bench_n$N.json
timeout 300 $L tools/train_bench.py --batch $((2048 / N)) --steps 20 --warmup 5 2> gpurun_out/train_n$N.err | tee gpurun_out/train_n${N}_strong.json | cut -c1-330; echo "train strong exit=$?"
timeout 300 $L tools/train_bench.py --batch 2048 --steps 10 --warmup 3 2>> gpurun_out/train_n$N.err | tee gpurun_out/train_n${N}_weak.json | cut -c1-330; echo "train weak exit=$?"
