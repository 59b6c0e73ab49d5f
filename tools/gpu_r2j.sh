#!/bin/bash
set -u
mkdir -p gpurun_out
# is the HBM-bound shape consumer-bound?  time per table pass at pass sizes 2 / 4 / 8 (Cfg<1,1> / <2,1> / <4,1>)
for n in 4800000 600000; do for g in 2 4 8; do
  timeout 300 python tools/run_step.py transe 256 $n 3 $g 2>&1 | tail -1 | cut -c1-200 | sed "s/^/N=$n group=$g: /"
done; done
timeout 300 python -m pytest tests/test_gpu_filter.py -q -k store_rows 2>&1 | tail -2
timeout 300 python tools/run_next_rows.py 2>&1 | grep -E "^f4"
