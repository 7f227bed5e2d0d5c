#!/bin/bash
# full ncu capture of the three tensor-core kernels (one launch each)
mkdir -p gpurun_out
timeout 300 python tools/prof_fwd.py --iters 2 --bwd > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|bwd_weight_tc_kernel|bwd_input_tc_kernel" -s 3 -c 3 -f -o gpurun_out/prof_tc \
    python tools/prof_fwd.py --iters 2 --bwd > gpurun_out/prof_ncu.log 2>&1
echo "exit=$?"; cat gpurun_out/prof_plain.log; tail -n 3 gpurun_out/prof_ncu.log
