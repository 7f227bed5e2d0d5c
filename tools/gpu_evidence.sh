#!/bin/bash
# round-end evidence: launch list of the bench command + full ncu captures of the three conv kernels
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit=$?"
timeout 300 python tools/prof_fwd.py --iters 2 --bwd > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|bwd_weight_tc_kernel|bwd_input_tc_kernel|bwd_alpha_partial" -s 4 -c 4 -f -o gpurun_out/prof_all \
    python tools/prof_fwd.py --iters 2 --bwd > gpurun_out/prof_ncu.log 2>&1
echo "ncu exit=$?"; tail -n 2 gpurun_out/prof_ncu.log
