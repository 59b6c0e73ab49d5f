#!/bin/bash
# end-of-round validation in the driver's own form + ncu evidence for the stream store kernel
set -u
mkdir -p gpurun_out/ncu5
echo "== pytest tests/ -x -q -m gpu"; timeout 1400 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_full.txt 2>&1; tail -2 gpurun_out/pytest_full.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash tools/gpu_bench.sh > gpurun_out/bench_run.txt 2>&1
grep -E "^(value|e2e|sharded|clocks)" gpurun_out/bench_run.txt | cut -c1-300
grep -E '"impl": "reference"' gpurun_out/bench_run.txt | cut -c1-200
timeout 300 ncu --set full --clock-control none -f -k regex:store_rows_stream_kernel -s 30 -c 1 -o gpurun_out/ncu5/store_rows_stream_4p8m python tools/run_store.py > gpurun_out/ncu5/l1.log 2>&1
for f in gpurun_out/ncu5/*.ncu-rep; do ncu -i $f --page raw --csv > ${f%.ncu-rep}.raw.csv 2>/dev/null; done
rm -f gpurun_out/ncu5/*.ncu-rep; ls gpurun_out/ncu5
