#!/bin/bash
# Fast-mode (tcgen05) bring-up: run its tests under a short timeout first.
set -u
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_fast.py -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_fast.txt
