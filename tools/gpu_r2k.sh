#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fast.py -q --tb=short -x 2>&1 | tail -25
