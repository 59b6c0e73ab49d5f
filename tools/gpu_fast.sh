#!/bin/bash
# Fast-mode (tcgen05) bring-up: run its tests under a short timeout first, then timings.
set -u
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_fast.py -x -q 2>&1 | tail -30 | tee gpurun_out/pytest_fast.txt
for spec in "distmult 1024 14541 20 fast" "complex 1024 40943 10 fast" "simple 1024 14541 20 fast" "distmult 8192 14541 10 fast" "distmult 64 4800000 3 fast" "distmult 64 14541 20 fast"; do
  timeout 120 python tools/run_sweep.py $spec 2>&1 | tail -1 | tee -a gpurun_out/sweeps_fast.txt
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fast_sweep_kernel -s 3 -c 1 -f -o gpurun_out/sweep_fast_distmult_fb \
  python tools/run_sweep.py distmult 1024 14541 2 fast > gpurun_out/ncu_fast.log 2>&1
