#!/bin/bash
# ResNet-20 w3a3 CiM data-parallel training at global batch 2048 (BASELINE.json config 4), CUDA-graphed steps
mkdir -p gpurun_out
: > gpurun_out/scale4.jsonl
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N"; fi
  timeout 200 $L tools/train_bench.py --batch $((2048 / N)) --steps 10 --warmup 3 >> gpurun_out/scale4.jsonl 2>> gpurun_out/scale4.err; echo "train N=$N exit=$?"
done
# weak scaling at 256 images per GPU
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/train_bench.py --batch 256 --steps 10 --warmup 3 >> gpurun_out/scale4.jsonl 2>> gpurun_out/scale4.err; echo "train weak N=8 exit=$?"
python - <<'PY'
import json
for l in open('gpurun_out/scale4.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    print(d['metric'], 'N=',d['n_gpus'], 'value=%.1f'%d['value'], d['unit'], 'ms/step=%.3f'%d['ms_per_step'], 'batch/gpu', d['config'].get('batch_per_gpu'))
PY
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/scale4.err | tail -n 4
