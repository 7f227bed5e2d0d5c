#!/bin/bash
for p in 8 16 32; do echo P=$p; CIMQ_ALPHA_P=$p timeout 300 python tools/time_bwd.py --only v2 2>&1 | grep "alpha"; done
