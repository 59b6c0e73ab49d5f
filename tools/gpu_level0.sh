#!/bin/bash
# integration levels (untouched train.py through patch() vs the fused sweep) on an FB15k-237-shaped evaluation
set -u
mkdir -p gpurun_out
rm -f gpurun_out/level0.txt gpurun_out/level0_profile.txt
env BLP_LEVEL0_OUT=gpurun_out/level0.txt ${PROFILE:+BLP_LEVEL0_PROFILE=gpurun_out/level0_profile.txt} timeout 800 python -m pytest tests/test_gpu_level0.py -q --tb=short -x ${PROFILE:+-k transe} 2>&1 | tail -15
cat gpurun_out/level0.txt
