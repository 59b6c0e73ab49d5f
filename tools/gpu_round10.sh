#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest"; timeout 600 python -m pytest tests/test_gpu_filter.py -x -q 2>&1 | tail -3
timeout 600 python tools/run_next_rows.py 2>&1 | grep "^f4" | tee gpurun_out/next_rows_f4.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:train_kernel -s 4 -c 2 -f -o gpurun_out/train_kernel python tools/run_train.py transe margin > /dev/null 2>&1
