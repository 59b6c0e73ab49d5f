#!/bin/bash
# full GPU parity pass + smoke + ncu evidence for the tensor-core sweep (plain and filter + refine)
set -u
mkdir -p gpurun_out/ncu2
echo "== pytest -m gpu"; timeout 1400 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/pytest_full.txt 2>&1; tail -3 gpurun_out/pytest_full.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
N="ncu --set full --clock-control none -f"
echo "== ncu fast sweep (plain) / fast_exact / refine"
timeout 300 $N -k regex:fast_sweep_kernel -s 3 -c 1 -o gpurun_out/ncu2/fast_sweep_distmult_16k python tools/run_sweep.py distmult 16384 14541 3 fast > gpurun_out/ncu2/l1.log 2>&1
timeout 300 $N -k regex:fast_sweep_kernel -s 3 -c 1 -o gpurun_out/ncu2/fast_exact_sweep_distmult_16k python tools/run_sweep.py distmult 16384 14541 3 fast_exact > gpurun_out/ncu2/l2.log 2>&1
timeout 300 $N -k regex:refine_kernel -s 3 -c 1 -o gpurun_out/ncu2/refine_distmult_16k python tools/run_sweep.py distmult 16384 14541 3 fast_exact > gpurun_out/ncu2/l3.log 2>&1
timeout 300 $N -k regex:fast_sweep_kernel -s 3 -c 1 -o gpurun_out/ncu2/fast_sweep_distmult_1k python tools/run_sweep.py distmult 1024 14541 3 fast > gpurun_out/ncu2/l4.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ncu2/launches_fast_exact.csv python tools/run_sweep.py distmult 16384 14541 2 fast_exact > /dev/null 2>&1
for f in gpurun_out/ncu2/*.ncu-rep; do
  ncu -i $f --page raw --csv > ${f%.ncu-rep}.raw.csv 2>/dev/null
done
rm -f gpurun_out/ncu2/*.ncu-rep
ls gpurun_out/ncu2
