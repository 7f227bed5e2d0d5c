#!/bin/bash
# SURVEY 8d configs 1/3/5 on one GPU: ResNet-20 w3a3 (batch 256 / 512 / 2048), w4a4 (196 its = one epoch of 256),
# wide (x4) w2a2 adc1 B=512
mkdir -p gpurun_out
: > gpurun_out/r02_train_1gpu.jsonl
run() { name=$1; shift; timeout 900 python tools/train_bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name exit=$?"; cut -c1-400 gpurun_out/$name.json; grep '^{' gpurun_out/$name.json >> gpurun_out/r02_train_1gpu.jsonl; tail -n 2 gpurun_out/$name.err; }
run train_w3a3 --batch 256 --steps 20 --warmup 5
run train_w3a3_b512 --batch 512 --steps 20 --warmup 5
run train_w3a3_b2048 --batch 2048 --steps 10 --warmup 3
run train_w4a4_epoch --batch 256 --nbits 4 --steps 196 --warmup 5
run train_wide_w2a2 --batch 512 --nbits 2 --adcbits 1 --width 4 --steps 10 --warmup 3
