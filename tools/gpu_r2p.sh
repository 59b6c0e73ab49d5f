#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_fullsize.py -q --tb=short -x 2>&1 | tail -4
for t in 1 0; do
  BLP_TRAIN_TMARED=$t timeout 120 python tools/run_train.py transe margin 2>&1 | tail -4 | sed "s/^/tmared=$t: /"
done
BLP_TRAIN_TMARED=1 timeout 120 python tools/run_train.py distmult margin 2>&1 | tail -4 | sed "s/^/tmared=1: /"
BLP_TRAIN_TMARED=0 timeout 120 python tools/run_train.py distmult margin 2>&1 | tail -4 | sed "s/^/tmared=0: /"
