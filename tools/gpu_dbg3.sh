#!/bin/bash
mkdir -p gpurun_out
for d in 0 31; do
CIMQ_V2_DBG=$d timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/l_dbg$d.csv python tools/time_bwd.py --only v2 --iters 1 > gpurun_out/l_dbg$d.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/l_dbg$d.csv')) if len(r)>5]
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value')
print('dbg=$d')
for r in rows[1:]:
    print('  ', r[ik][:70], r[iv])
PY
done
