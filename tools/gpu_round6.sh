#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
echo "== host overhead"; timeout 300 python tools/host_overhead.py 2>&1 | grep "us / call" | tee gpurun_out/host_overhead.txt
echo "== bench exact"; timeout 600 python bench.py --steps 50 --warmup 5 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-250; tail -3 gpurun_out/bench.err
echo "== bench fast distmult"; timeout 600 python bench.py --steps 50 --warmup 5 --model distmult --mode fast --no-extra --no-cpu-baseline 2> gpurun_out/bench_fast.err | tee gpurun_out/bench_fast_distmult.json | cut -c1-250; tail -3 gpurun_out/bench_fast.err
