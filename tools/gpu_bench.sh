#!/bin/bash
# bench + launch list + one full ncu capture of the dominant kernels (development helper)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "full_size" > gpurun_out/t7_full.log 2>&1; echo "t7 exit=$?"; tail -n 3 gpurun_out/t7_full.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit=$?"
