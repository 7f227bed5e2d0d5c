#!/bin/bash
# per-role cycle counters of the wgrad kernel (library built with `make TIMERS=1`) at a layer geometry
CH=${CH:-16}; HW=${HW:-32}; B=${B:-1024}
for d in ${DBGS:-0 63}; do
  echo "--- CIMQ_V2_DBG_WG=$d"
  CIMQ_V2_DBG_WG=$d python tools/prof_v2.py --channels $CH --hw $HW --batch $B --iters 3 --time --timers 2>&1 | grep -E "forward|timer|wgrad block"
done
