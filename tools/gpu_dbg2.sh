#!/bin/bash
timeout 300 python tools/time_fwd.py 2>&1 | grep "avg\|diff"
