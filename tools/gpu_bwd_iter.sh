#!/bin/bash
# backward iteration: v2 parity tests, then device timing of the backward parts
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?"; tail -n ${TAILN:-12} gpurun_out/$name.log; }
run i_v2 python -m pytest tests/test_gpu_v2.py -m gpu -q -x
run i_matrix python -m pytest tests/test_gpu_matrix.py -m gpu -q -x
run i_time python tools/time_bwd.py
