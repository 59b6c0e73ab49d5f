#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_step.py tests/test_gpu_eval.py -q --tb=short 2>&1 | tail -3
BLP_SWEEP_CFG=5 timeout 600 python -m pytest tests/test_gpu_step.py tests/test_gpu_eval.py tests/test_gpu_fullsize.py -q --tb=short 2>&1 | tail -3
for cfg in 4 5; do for m in transe distmult complex; do
  SORT_REL=1 BLP_SWEEP_CFG=$cfg timeout 120 python tools/run_step.py $m 1024 14541 30 2>&1 | tail -1 | cut -c1-140 | sed "s/^/cfg$cfg sorted: /"
done; BLP_SWEEP_CFG=$cfg timeout 120 python tools/run_step.py transe 1024 14541 30 2>&1 | tail -1 | cut -c1-140 | sed "s/^/cfg$cfg: /"; done
