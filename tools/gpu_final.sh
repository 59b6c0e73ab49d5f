#!/bin/bash
# end-of-round validation: parity suite, smoke, bench (both arms)
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1400 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_full.txt 2>&1; tail -2 gpurun_out/pytest_full.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash tools/gpu_bench.sh > gpurun_out/bench_run.txt 2>&1
grep -E "^(value|e2e|sharded|clocks)" gpurun_out/bench_run.txt | cut -c1-300
grep -E '"impl": "reference"' gpurun_out/bench_run.txt | cut -c1-200
