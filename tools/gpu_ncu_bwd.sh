#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_fwd.py --iters 2 --bwd > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"bwd_weight_tc_kernel|bwd_input_tc_kernel" -s 2 -c 2 -f -o gpurun_out/prof_bwd \
    python tools/prof_fwd.py --iters 2 --bwd > gpurun_out/prof_ncu.log 2>&1
echo "exit=$?"; cat gpurun_out/prof_plain.log; tail -n 2 gpurun_out/prof_ncu.log
