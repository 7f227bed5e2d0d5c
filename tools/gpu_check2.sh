#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?"; tail -n 4 gpurun_out/$name.log; }
run t2_cabi_tc  python -m pytest tests -m gpu -q -x -k "cabi and False"
run t6_random   python -m pytest tests -m gpu -q -x -k "random_layer"
run t7_full     python -m pytest tests -m gpu -q -x -k "full_size"
run t9_all      python -m pytest tests -m gpu -q
run b1_prof     python tools/prof_fwd.py --iters 3 --bwd
run b2_bench    python bench.py --steps 10 --warmup 3
