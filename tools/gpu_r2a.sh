#!/bin/bash
# round-2 first pass: parity of the fused step + per-call timings.  gpurun --timeout 1500 -- 'bash tools/gpu_r2a.sh'
set -u
mkdir -p gpurun_out
echo "== pytest step"; timeout 900 python -m pytest tests/test_gpu_step.py -x -q 2>&1 | tail -15 | tee gpurun_out/r2a_pytest_step.txt
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/r2a_pytest_gpu.txt
echo "== step timings"
rm -f gpurun_out/r2a_steps.txt
for spec in "transe 64 14541 300" "transe 64 14541 300 8" "transe 64 14541 300 32" "distmult 64 14541 300" "transe 1024 14541 50" "complex 64 40943 100" "transe 128 14541 200" "transe 256 14541 100"; do
  timeout 300 python tools/run_step.py $spec 2>&1 | tail -1 | tee -a gpurun_out/r2a_steps.txt
done
SORT_REL=1 timeout 300 python tools/run_step.py transe 1024 14541 50 2>&1 | tail -1 | tee -a gpurun_out/r2a_steps.txt
timeout 300 python tools/run_step.py transe 64 4800000 5 2 2>&1 | tail -1 | tee -a gpurun_out/r2a_steps.txt
echo "== host overhead"; timeout 300 python tools/host_overhead.py 2>&1 | grep "us / call" | tee gpurun_out/r2a_host_overhead.txt
