#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q --tb=short 2>&1 > gpurun_out/r2b_pytest_full.txt; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2b_pytest_full.txt | sed -E 's/ - .*//' | head -40
grep -E "^E  " gpurun_out/r2b_pytest_full.txt | sort | uniq -c | sort -rn | head -20
echo "== phases"
for spec in "transe 64 14541" "transe 1024 14541" "complex 64 14541"; do
  echo "--- $spec"; timeout 300 python tools/step_phases.py $spec 2>&1 | tail -11
done | tee gpurun_out/r2b_phases.txt
echo "== step timings"
for spec in "transe 64 14541 300" "distmult 64 14541 300" "transe 1024 14541 50" "transe 256 14541 100"; do
  timeout 300 python tools/run_step.py $spec 2>&1 | tail -1 | tee -a gpurun_out/r2b_steps.txt
done
SORT_REL=1 timeout 300 python tools/run_step.py transe 1024 14541 50 2>&1 | tail -1 | tee -a gpurun_out/r2b_steps.txt
