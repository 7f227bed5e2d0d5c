#!/bin/bash
# f-1: the reference's own main_lsq.py, unchanged, on our kernels and on its own PyTorch-CUDA modules
mkdir -p gpurun_out
timeout 900 python -m cim_quantization_b200.launcher --impl ours --train-batches 16 --val-batches 2 --epochs 2 > gpurun_out/launcher_ours.log 2>&1; echo "ours exit=$?"
grep -E "LAUNCHER_RESULT|Acc@1|Error|error" gpurun_out/launcher_ours.log | tail -8
timeout 1500 python -m cim_quantization_b200.launcher --impl reference --train-batches 16 --val-batches 2 --epochs 2 > gpurun_out/launcher_ref.log 2>&1; echo "ref exit=$?"
grep -E "LAUNCHER_RESULT|Acc@1|Error|error" gpurun_out/launcher_ref.log | tail -8
