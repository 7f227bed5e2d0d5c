#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bwd_weight_tc_kernel|bwd_input_v2|conv_v2_kernel" -s 2 -c 3 -o gpurun_out/r02_small_l3 -f python tools/time_bwd.py --only v2 --cin 64 --cout 64 --hw 8 --iters 1 > gpurun_out/ncu_small.log 2>&1
echo "exit=$?"; tail -3 gpurun_out/ncu_small.log
