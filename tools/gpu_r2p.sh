#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_fullsize.py -q --tb=short -x 2>&1 | tail -3
timeout 120 python tools/run_train.py transe margin 2>&1 | tail -4
timeout 120 python tools/run_train.py transe nll 2>&1 | tail -4
