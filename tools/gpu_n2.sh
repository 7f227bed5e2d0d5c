#!/bin/bash
# two-rank sanity of the driver's multi-GPU launch: bench.py both arms + the ResNet training bench
mkdir -p gpurun_out
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 300 $L bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench N=2 exit=$?"; cat gpurun_out/bench_n2.json | cut -c1-600
timeout 300 $L bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref N=2 exit=$?"; cat gpurun_out/bench_ref_n2.json | cut -c1-300
timeout 300 $L tools/train_bench.py --batch 256 --steps 10 --warmup 3 2> gpurun_out/train_n2.err | cut -c1-300; echo "train N=2 exit=$?"
