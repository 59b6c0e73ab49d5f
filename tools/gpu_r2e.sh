#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1400 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_full.txt 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_full.txt | sed -E 's/ - .*//' | head -30
grep -E "^E  " gpurun_out/pytest_full.txt | sort | uniq -c | sort -rn | head -8 | cut -c1-300
SORT_REL=1 timeout 120 python tools/run_step.py transe 1024 14541 30 2>&1 | tail -1
timeout 120 python tools/run_step.py transe 1024 14541 30 2>&1 | tail -1
bash tools/gpu_bench.sh
