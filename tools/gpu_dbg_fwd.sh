#!/bin/bash
# forward role isolation at a layer geometry (CIMQ_V2_DBG: 1 no GEMM2, 2 no GEMM1, 4 no epilogue arithmetic, 8 no row assembly, 16 one producer group)
CH=${CH:-16}; HW=${HW:-32}; B=${B:-1024}
for d in 0 1 2 4 8 12 16; do
  echo -n "dbg=$d  "; CIMQ_V2_DBG=$d python tools/prof_v2.py --channels $CH --hw $HW --batch $B --iters 4 --time 2>&1 | grep forward
done
