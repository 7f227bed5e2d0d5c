#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_v2.py tests/test_gpu_matrix.py -m gpu -q -x 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"go_scales|alpha_finish|bwd_input_v2|bwd_alpha_v3" -s 6 -c 8 --csv --log-file gpurun_out/l_small.csv python tools/prof_v2.py --iters 4 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/l_small.csv
timeout 300 python tools/time_bwd.py --only v2 2>&1 | grep -v "^grad"
