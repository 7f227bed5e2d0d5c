#!/bin/bash
# ResNet-20 w3a3 CiM data-parallel training at global batch 2048 (BASELINE.json config 4), CUDA-graphed steps
mkdir -p gpurun_out
: > gpurun_out/scale3.jsonl
timeout 300 python -m pytest tests -m gpu -q -x -k "lsq_matches or act_lsq or linear_lsq or conv2d_lsq" > gpurun_out/t_lsqmod.log 2>&1; echo "tests exit=$?"; tail -n 3 gpurun_out/t_lsqmod.log
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N"; fi
  timeout 200 $L tools/train_bench.py --batch $((2048 / N)) --steps 10 --warmup 3 >> gpurun_out/scale3.jsonl 2>> gpurun_out/scale3.err; echo "train N=$N exit=$?"
done
python - <<'PY'
import json
for l in open('gpurun_out/scale3.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    print(d['metric'], 'N=',d['n_gpus'], 'value=%.1f'%d['value'], d['unit'], 'ms/step=%.3f'%d['ms_per_step'], d['config'].get('cuda_graph'))
PY
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/scale3.err | tail -n 4
