#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fast.py -q --tb=short -x 2>&1 | tail -3
for dbg in 0 1 4; do
  BLP_FAST_DEBUG=$dbg timeout 120 python tools/run_sweep.py distmult 16384 14541 5 fast 2>&1 | tail -1 | cut -c1-120 | sed "s/^/debug=$dbg: /"
done
timeout 120 python tools/run_sweep.py distmult 16384 14541 5 fast_exact 2>&1 | tail -1
timeout 120 python tools/run_sweep.py complex 3136 40943 5 fast 2>&1 | tail -1
timeout 120 python tools/run_sweep.py complex 3136 40943 5 fast_exact 2>&1 | tail -1
timeout 120 python tools/run_sweep.py distmult 1024 14541 20 fast 2>&1 | tail -1
