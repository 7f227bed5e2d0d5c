#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "launcher or deterministic" > gpurun_out/t_new.log 2>&1; echo "exit=$?"; tail -n 5 gpurun_out/t_new.log
