#!/bin/bash
# wgrad role isolation at a given layer geometry: backward time of the layer with parts of the wgrad kernel switched off
CH=${CH:-16}; HW=${HW:-32}; B=${B:-1024}
for d in 0 1 2 8 10 4 32 16 63; do
  echo -n "dbg=$d  "; CIMQ_V2_DBG_WG=$d python tools/prof_v2.py --channels $CH --hw $HW --batch $B --iters 4 --time 2>&1 | grep forward
done
