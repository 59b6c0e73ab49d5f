#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
rm -f gpurun_out/sweeps.txt
for spec in "transe 1024 14541 20" "transe 2 4800000 10" "transe 64 4800000 3"; do
  timeout 120 python tools/run_sweep.py $spec 2>&1 | tail -1 | tee -a gpurun_out/sweeps.txt
  SORT_REL=1 timeout 120 python tools/run_sweep.py $spec 2>&1 | tail -1 | sed 's/^/sorted: /' | tee -a gpurun_out/sweeps.txt
done
echo "== bench exact"; timeout 600 python bench.py --steps 50 --warmup 5 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-250; tail -3 gpurun_out/bench.err
