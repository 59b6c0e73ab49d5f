#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lazy.py tests/test_gpu_filter.py tests/test_gpu_eval.py -q --tb=short -x 2>&1 | tail -3
bash tools/gpu_level0.sh
