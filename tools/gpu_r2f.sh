#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1400 python -m pytest tests/test_gpu_train.py tests/test_gpu_fullsize.py -q --tb=short > gpurun_out/pytest_full.txt 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_full.txt | sed -E 's/ - .*//' | head -30
grep -E "^E  " gpurun_out/pytest_full.txt | sort | uniq -c | sort -rn | head -8 | cut -c1-300
for m in "transe margin" "distmult margin" "complex nll" "simple margin"; do timeout 300 python tools/run_train.py $m 2>&1 | tail -4 | cut -c1-120; done | tee gpurun_out/train_kernel.txt
