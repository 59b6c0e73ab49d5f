#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.txt
rm -f gpurun_out/sweeps.txt
for spec in "transe 1024 14541 20" "distmult 1024 14541 20 fast" "complex 1024 40943 10 fast" "distmult 8192 14541 10 fast" "distmult 64 14541 20 fast" "distmult 64 4800000 3 fast"; do
  timeout 120 python tools/run_sweep.py $spec 2>&1 | tail -1 | tee -a gpurun_out/sweeps.txt
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_fast_1024.csv \
  python tools/run_sweep.py distmult 1024 14541 3 fast > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/launches_fast_1024.csv
echo "== bench fast distmult"; timeout 600 python bench.py --steps 50 --warmup 5 --model distmult --mode fast --no-extra --no-cpu-baseline 2> gpurun_out/bench_fast.err | tee gpurun_out/bench_fast_distmult.json | cut -c1-300; tail -3 gpurun_out/bench_fast.err
