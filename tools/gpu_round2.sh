#!/bin/bash
# new-row tests + bench in both modes.  gpurun --timeout 1200 -- 'bash tools/gpu_round2.sh'
set -u
mkdir -p gpurun_out
echo "== pytest new"; timeout 600 python -m pytest tests/test_gpu_filter.py -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_filter.txt
echo "== bench exact"; timeout 600 python bench.py --steps 50 --warmup 5 2> gpurun_out/bench.err | tee gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== bench fast distmult"; timeout 600 python bench.py --steps 50 --warmup 5 --model distmult --mode fast --no-extra 2> gpurun_out/bench_fast.err | tee gpurun_out/bench_fast_distmult.json; tail -3 gpurun_out/bench_fast.err
