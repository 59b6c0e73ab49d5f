#!/bin/bash
set -u
timeout 300 python -m pytest tests/test_gpu_fast.py -q --tb=short -x 2>&1 | tail -2
BLP_FAST_PAIR=0 timeout 300 python -m pytest tests/test_gpu_fast.py -q --tb=short -x 2>&1 | tail -2
timeout 120 python tools/run_sweep.py distmult 16384 14541 5 fast 2>&1 | tail -1 | cut -c1-130
BLP_FAST_DEBUG=4 timeout 120 python tools/run_sweep.py distmult 16384 14541 5 fast 2>&1 | tail -1 | cut -c1-130 | sed "s/^/epi-only: /"
BLP_FAST_DEBUG=1 timeout 120 python tools/run_sweep.py distmult 16384 14541 5 fast 2>&1 | tail -1 | cut -c1-130 | sed "s/^/mma-only: /"
timeout 120 python tools/run_sweep.py distmult 16384 14541 5 fast_exact 2>&1 | tail -1 | cut -c1-130
timeout 120 python tools/run_sweep.py complex 3136 40943 5 fast_exact 2>&1 | tail -1 | cut -c1-130
timeout 120 python tools/run_sweep.py distmult 1024 14541 20 fast 2>&1 | tail -1 | cut -c1-130
