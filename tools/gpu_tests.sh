#!/bin/bash
# parity pass: gpurun --timeout 1500 -- 'bash tools/gpu_tests.sh [pytest args]'
set -u
mkdir -p gpurun_out
timeout 1400 python -m pytest tests -m gpu -q --tb=short "$@" > gpurun_out/pytest_full.txt 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_full.txt | sed -E 's/ - .*//' | head -40
grep -E "^E  " gpurun_out/pytest_full.txt | sort | uniq -c | sort -rn | head -12 | cut -c1-300
