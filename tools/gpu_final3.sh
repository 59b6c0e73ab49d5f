#!/bin/bash
# end-of-round validation in the driver's own form
set -u
mkdir -p gpurun_out
echo "== pytest tests/ -x -q -m gpu"; timeout 1400 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_full.txt 2>&1; tail -2 gpurun_out/pytest_full.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash tools/gpu_bench.sh > gpurun_out/bench_run.txt 2>&1
grep -E "^(value|e2e|sharded|clocks)" gpurun_out/bench_run.txt | cut -c1-300
grep -E "^leg store" gpurun_out/bench_run.txt | cut -c1-400
grep -E '"impl": "reference"' gpurun_out/bench_run.txt | cut -c1-200
