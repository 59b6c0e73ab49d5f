#!/bin/bash
set -u
mkdir -p gpurun_out/ncu2
timeout 300 ncu --set full --clock-control none --import-source on -f -k regex:train128_kernel -s 30 -c 1 -o gpurun_out/ncu2/train128_b1024 python tools/run_train.py transe margin > gpurun_out/ncu2/l5.log 2>&1
ncu -i gpurun_out/ncu2/train128_b1024.ncu-rep --page raw --csv > gpurun_out/ncu2/train128_b1024.raw.csv 2>/dev/null
ncu -i gpurun_out/ncu2/train128_b1024.ncu-rep --page source --csv > gpurun_out/ncu2/train128_b1024.source.csv 2>/dev/null
rm -f gpurun_out/ncu2/train128_b1024.ncu-rep
ls -la gpurun_out/ncu2 | tail -5
