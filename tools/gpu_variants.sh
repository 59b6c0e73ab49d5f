#!/bin/bash
# time the tuning variants under blp_b200/variants/: gpurun --timeout 900 -- 'bash tools/gpu_variants.sh'
set -u
mkdir -p gpurun_out
for so in "" blp_b200/variants/*.so; do
  echo "--- ${so:-default}"
  for spec in "transe 1024 14541 30"; do
    BLP_B200_LIB=${so:+$PWD/$so} timeout 120 python tools/run_step.py $spec 2>&1 | tail -1 | cut -c1-150
    SORT_REL=1 BLP_B200_LIB=${so:+$PWD/$so} timeout 120 python tools/run_step.py $spec 2>&1 | tail -1 | cut -c1-150 | sed 's/^/   sorted: /'
  done
done | tee gpurun_out/variants.txt
