#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_fwd.py --iters 2 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 1 -f -o gpurun_out/prof_fwd \
    python tools/prof_fwd.py --iters 2 > gpurun_out/prof_ncu.log 2>&1
echo "exit=$?"; cat gpurun_out/prof_plain.log; tail -n 3 gpurun_out/prof_ncu.log
