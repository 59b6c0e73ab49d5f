#!/bin/bash
# warp-per-triple fold kernel + vectorised metrics: parity, smoke, launch list of a whole fast call, bench
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1400 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_full.txt 2>&1; tail -2 gpurun_out/pytest_full.txt
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_full.txt | head -20
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for m in fast fast_exact; do timeout 120 python tools/run_plan.py distmult 20480 14541 20 $m 2>&1 | tail -1; done
timeout 120 python tools/run_plan.py complex 3136 40943 20 fast 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fast_plan.csv python tools/run_plan.py distmult 20480 14541 2 fast > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/launches_fast_plan.csv | head -12
bash tools/gpu_bench.sh > gpurun_out/bench_run.txt 2>&1
grep -E "^(value|e2e|sharded|clocks)" gpurun_out/bench_run.txt | cut -c1-300
grep -E "^leg " gpurun_out/bench_run.txt | cut -c1-260
grep -E '"impl": "reference"' gpurun_out/bench_run.txt | cut -c1-200
