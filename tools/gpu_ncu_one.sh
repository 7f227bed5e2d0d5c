#!/bin/bash
# usage: gpu_ncu_one.sh <kernel-regex> [prof_fwd.py args...]
mkdir -p gpurun_out
K=$1; shift
timeout 300 python tools/prof_fwd.py --iters 2 "$@" > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$K" -s 1 -c 1 -f -o gpurun_out/prof_one \
    python tools/prof_fwd.py --iters 2 "$@" > gpurun_out/prof_ncu.log 2>&1
echo "exit=$?"; cat gpurun_out/prof_plain.log; tail -n 2 gpurun_out/prof_ncu.log
