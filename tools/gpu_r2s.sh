#!/bin/bash
set -u
BLP_SWEEP_CFG=6 timeout 600 python -m pytest tests/test_gpu_step.py tests/test_gpu_eval.py tests/test_gpu_fullsize.py -q --tb=short -x 2>&1 | tail -2
for n in 4800000 600000; do for cfg in 1 6; do
  BLP_SWEEP_CFG=$cfg timeout 300 python tools/run_step.py transe 256 $n 3 0 2>&1 | tail -1 | cut -c1-150 | sed "s/^/N=$n cfg=$cfg: /"
done; done
for cfg in 1 6 5; do for m in distmult complex; do
  BLP_SWEEP_CFG=$cfg timeout 300 python tools/run_step.py $m 256 600000 3 0 2>&1 | tail -1 | cut -c1-150 | sed "s/^/N=600000 cfg=$cfg: /"
done; done
