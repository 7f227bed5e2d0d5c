#!/bin/bash
# both bench arms as the driver runs them (development helper)
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit=$?"; cat gpurun_out/bench_ref.json; tail -n 5 gpurun_out/bench_ref.err
