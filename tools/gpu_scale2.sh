#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/scale2.jsonl
timeout 300 python -m pytest tests -m gpu -q -x -k "alpha_quantizer or module_matches or resnet20" > gpurun_out/t_aq.log 2>&1; echo "tests exit=$?"; tail -n 3 gpurun_out/t_aq.log
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 300 python tools/train_bench.py --batch 256 --steps 10 --warmup 3 >> gpurun_out/scale2.jsonl 2>> gpurun_out/scale2.err; echo "train N=1 exit=$?"
timeout 300 $L bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline >> gpurun_out/scale2.jsonl 2>> gpurun_out/scale2.err; echo "bench N=2 exit=$?"
timeout 300 $L tools/train_bench.py --batch 256 --steps 10 --warmup 3 >> gpurun_out/scale2.jsonl 2>> gpurun_out/scale2.err; echo "train N=2 exit=$?"
python - <<'PY'
import json
for l in open('gpurun_out/scale2.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    print(d['metric'], 'N=',d['n_gpus'], 'value=%.1f'%d['value'], d['unit'], 'ms/step=%.3f'%d['ms_per_step'], d['config'].get('cuda_graph'))
PY
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/scale2.err | tail -n 6
