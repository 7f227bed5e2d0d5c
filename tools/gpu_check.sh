#!/bin/bash
# Staged GPU check used during development: every stage runs in its own process under `timeout`
# so that a faulting kernel cannot take the later stages (or the box) with it.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?"; tail -n 6 gpurun_out/$name.log; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
run t1_quant    python -m pytest tests -m gpu -q -k "lsq or step_sizes or psums or alpha_cim_init"
run t2_simt     python -m pytest tests -m gpu -q -k "cabi and True"
run t3_module_s python -m pytest tests -m gpu -q -k "(module_matches and True) or function_17 or lazy_init"
run t4_tc       python -m pytest tests -m gpu -q -k "cabi and False"
run t5_module_t python -m pytest tests -m gpu -q -k "module_matches and False"
run t6_random   python -m pytest tests -m gpu -q -k "random_layer"
run t7_full     python -m pytest tests -m gpu -q -k "full_size"
run t8_smoke    python __graft_entry__.py smoke
