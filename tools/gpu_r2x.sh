#!/bin/bash
# vectorised metrics kernel: parity, launch list of a whole fast call, ncu of fold_triples / metrics
set -u
mkdir -p gpurun_out/ncu3
echo "== pytest -m gpu"; timeout 1400 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_full.txt 2>&1; tail -2 gpurun_out/pytest_full.txt
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_full.txt | head -20
for m in fast fast_exact; do timeout 120 python tools/run_plan.py distmult 20480 14541 20 $m 2>&1 | tail -1; done
timeout 120 python tools/run_plan.py complex 3136 40943 20 fast 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fast_plan.csv python tools/run_plan.py distmult 20480 14541 2 fast > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/launches_fast_plan.csv | head -12
N="ncu --set full --clock-control none -f"
timeout 300 $N -k regex:fold_triples_kernel -s 3 -c 1 -o gpurun_out/ncu3/fold_triples_20k python tools/run_plan.py distmult 20480 14541 2 fast > gpurun_out/ncu3/l1.log 2>&1
timeout 300 $N -k regex:metrics_reduce_cluster_kernel -s 3 -c 1 -o gpurun_out/ncu3/metrics_reduce_40k python tools/run_plan.py distmult 20480 14541 2 fast > gpurun_out/ncu3/l2.log 2>&1
for f in gpurun_out/ncu3/*.ncu-rep; do ncu -i $f --page raw --csv > ${f%.ncu-rep}.raw.csv 2>/dev/null; done
rm -f gpurun_out/ncu3/*.ncu-rep
ls gpurun_out/ncu3
