#!/bin/bash
# final-code ncu evidence for the bench command: launch list + one --set full capture of the headline sweep launch
set -u
mkdir -p gpurun_out/ncu4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/ncu4/launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu4/bench_under_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/ncu4/launches_bench.csv | head -16
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:sweep_kernel -s 3 -c 1 -o gpurun_out/ncu4/sweep_transe_fb_e20480 \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu4/l1.log 2>&1
tail -3 gpurun_out/ncu4/l1.log | cut -c1-200
for f in gpurun_out/ncu4/*.ncu-rep; do
  ncu -i $f --page raw --csv > ${f%.ncu-rep}.raw.csv 2>/dev/null
  ncu -i $f --page details --csv > ${f%.ncu-rep}.details.csv 2>/dev/null
done
rm -f gpurun_out/ncu4/*.ncu-rep
ls -la gpurun_out/ncu4
