#!/bin/bash
set -u
mkdir -p gpurun_out
for E in 1024 8192; do
timeout 300 ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_elapsed.max --clock-control none -c 40 --csv --log-file gpurun_out/launches_fast_$E.csv \
  python tools/run_sweep.py distmult $E 14541 3 fast > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/launches_fast_$E.csv
for dbg in 5; do
  BLP_FAST_DEBUG=$dbg timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_fast_${E}_dbg$dbg.csv \
    python tools/run_sweep.py distmult $E 14541 3 fast > /dev/null 2>&1
  python tools/launch_summary.py gpurun_out/launches_fast_${E}_dbg$dbg.csv
done
done
