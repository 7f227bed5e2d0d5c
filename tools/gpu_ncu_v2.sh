#!/bin/bash
# ncu --set full of one forward + backward of the microbench layer on the v2 kernels (after a plain run exits 0)
mkdir -p gpurun_out
timeout 300 python tools/prof_v2.py --iters 2 > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail gpurun_out/plain.log; exit 1; }
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-conv_v2|bwd_|go_scales}" -s ${SKIP:-4} -c ${COUNT:-5} \
  -o gpurun_out/${1:-r02_v2_b} -f python tools/prof_v2.py --iters 2 > gpurun_out/ncu.log 2>&1
echo "ncu exit=$?"; tail -5 gpurun_out/ncu.log
