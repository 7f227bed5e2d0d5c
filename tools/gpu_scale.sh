#!/bin/bash
# multi-GPU scaling runs: layer microbench (weak scaling) and ResNet-20 data-parallel training (global batch 2048)
mkdir -p gpurun_out
NG=${1:-8}
: > gpurun_out/scale.jsonl
for N in 1 2 4 8; do
  [ $N -gt $NG ] && break
  if [ $N -eq 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"; fi
  timeout 600 $L bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline >> gpurun_out/scale.jsonl 2>> gpurun_out/scale.err; echo "bench N=$N exit=$?"
  timeout 600 $L tools/train_bench.py --batch $((2048 / N)) --steps 10 --warmup 3 >> gpurun_out/scale.jsonl 2>> gpurun_out/scale.err; echo "train N=$N exit=$?"
done
python - <<'PY'
import json
for l in open('gpurun_out/scale.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    print(d['metric'], 'N=',d['n_gpus'], 'value=%.1f'%d['value'], d['unit'], 'ms/step=%.3f'%d['ms_per_step'])
PY
tail -n 5 gpurun_out/scale.err
