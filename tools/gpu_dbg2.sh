#!/bin/bash
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"finish|go_scales" -s 6 -c 9 --csv --log-file gpurun_out/l_small.csv python tools/prof_v2.py --iters 4 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/l_small.csv
timeout 900 python -m pytest tests/test_gpu_v2.py tests/test_gpu_matrix.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
