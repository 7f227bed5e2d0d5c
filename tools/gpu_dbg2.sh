#!/bin/bash
for d in 0 63; do echo "dbg=$d"; CIMQ_V2_DBG=$d timeout 300 python tools/time_bwd.py --only v2 2>&1 | grep "wgrad\|dgrad"; done
timeout 300 python tools/time_fwd.py 2>&1 | grep "v2_"
