#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fast.py -q --tb=short -x 2>&1 | tail -3
BLP_FAST_PAIR=0 timeout 300 python -m pytest tests/test_gpu_fast.py -q --tb=short -x 2>&1 | tail -3
for p in 0 1; do
  BLP_FAST_PAIR=$p timeout 120 python tools/run_sweep.py distmult 16384 14541 5 fast 2>&1 | tail -1 | cut -c1-130 | sed "s/^/pair=$p: /"
  BLP_FAST_PAIR=$p BLP_FAST_DEBUG=4 timeout 120 python tools/run_sweep.py distmult 16384 14541 5 fast 2>&1 | tail -1 | cut -c1-130 | sed "s/^/pair=$p epi-only: /"
  BLP_FAST_PAIR=$p timeout 120 python tools/run_sweep.py distmult 16384 14541 5 fast_exact 2>&1 | tail -1 | cut -c1-130 | sed "s/^/pair=$p: /"
  BLP_FAST_PAIR=$p timeout 120 python tools/run_sweep.py complex 3136 40943 5 fast_exact 2>&1 | tail -1 | cut -c1-130 | sed "s/^/pair=$p: /"
  BLP_FAST_PAIR=$p timeout 120 python tools/run_sweep.py distmult 1024 14541 20 fast 2>&1 | tail -1 | cut -c1-130 | sed "s/^/pair=$p: /"
done
