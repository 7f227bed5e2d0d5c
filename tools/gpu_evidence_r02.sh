#!/bin/bash
# Round-2 evidence: ncu --set full of every kernel of one microbench step on the v2 kernels (after a plain run exits 0),
# then the launch list of the default bench command (gpu__time_duration only).
mkdir -p gpurun_out
timeout 300 python tools/prof_v2.py --iters 2 > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail gpurun_out/plain.log; exit 1; }
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"conv_v2|bwd_|go_scales" -s 7 -c 7 \
  -o gpurun_out/r02_final -f python tools/prof_v2.py --iters 3 > gpurun_out/ncu.log 2>&1
echo "ncu exit=$?"; tail -3 gpurun_out/ncu.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-train --no-matrix --no-cpu-baseline > gpurun_out/b_plain.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-train --no-matrix --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "launch list exit=$?"; wc -l gpurun_out/r02_launches_bench.csv
