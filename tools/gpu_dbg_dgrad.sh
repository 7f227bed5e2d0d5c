#!/bin/bash
# which role bounds the v2 dgrad: time it with parts switched off (results are garbage then)
for d in 0 1 2 4 3 5 6 7; do echo "dbg=$d"; CIMQ_V2_DBG=$d timeout 300 python tools/time_bwd.py --only v2 2>&1 | grep "dgrad"; done
