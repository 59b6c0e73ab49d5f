#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_gpu_step.py tests/test_gpu_eval.py tests/test_gpu_lazy.py -q --tb=short -x 2>&1 | tail -3
for ov in "" 1; do
  OVERLAP=$ov timeout 120 python tools/run_step.py transe 64 14541 300 2>&1 | tail -1 | cut -c1-150 | sed "s/^/overlap=$ov: /"
  OVERLAP=$ov timeout 120 python tools/run_step.py distmult 64 14541 300 2>&1 | tail -1 | cut -c1-150 | sed "s/^/overlap=$ov: /"
  OVERLAP=$ov timeout 120 python tools/run_step.py transe 1024 14541 50 2>&1 | tail -1 | cut -c1-150 | sed "s/^/overlap=$ov: /"
done
