#!/bin/bash
set -u
mkdir -p gpurun_out
for so in "" blp_b200/variants/*.so; do
  echo "--- ${so:-default}"
  BLP_B200_LIB=${so:+$PWD/$so} timeout 120 python tools/run_train.py transe margin 2>&1 | tail -4 | cut -c1-110
done | tee gpurun_out/variants_train.txt
