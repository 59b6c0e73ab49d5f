#!/bin/bash
# ncu evidence pass (one GPU): gpurun --timeout 1700 -- 'bash tools/gpu_ncu.sh'
# launch list of the bench command + one `--set full` capture per kernel; summaries are made on the CPU box with
# tools/ncu_summary.py and committed under profiles/.
set -u
mkdir -p gpurun_out/ncu
N="ncu --set full --clock-control none -f"
NS="ncu --set full --clock-control none --import-source on -f"
echo "== launch list of the bench command"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/ncu/launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu/bench_under_ncu.log 2>&1
echo "== sweep transe FB (bench command)"
timeout 600 $NS -k regex:sweep_kernel -s 70 -c 1 -o gpurun_out/ncu/sweep_transe_fb python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu/l1.log 2>&1
echo "== sweep transe FB E=64 (one-launch step)"
timeout 300 $N -k regex:sweep_kernel -s 6 -c 1 -o gpurun_out/ncu/sweep_transe_fb_e64 python tools/run_step.py transe 64 14541 10 > gpurun_out/ncu/l2.log 2>&1
echo "== sweep distmult FB exact / complex WN exact"
timeout 300 $N -k regex:sweep_kernel -s 6 -c 1 -o gpurun_out/ncu/sweep_distmult_fb python tools/run_step.py distmult 1024 14541 5 > gpurun_out/ncu/l3.log 2>&1
timeout 300 $N -k regex:sweep_kernel -s 6 -c 1 -o gpurun_out/ncu/sweep_complex_wn python tools/run_step.py complex 1024 40943 3 > gpurun_out/ncu/l4.log 2>&1
echo "== sweep transe WD-scale 600k-row shard, table pass 2 (HBM-bound)"
timeout 300 $N -k regex:sweep_kernel -s 3 -c 1 -o gpurun_out/ncu/sweep_transe_wd600k python tools/run_step.py transe 64 600000 3 2 > gpurun_out/ncu/l5.log 2>&1
echo "== wide transe d=768"
timeout 300 $N -k regex:sweep_wide_kernel -s 2 -c 1 -o gpurun_out/ncu/sweep_wide_d768 python tools/run_wide.py 768 1024 > gpurun_out/ncu/l6.log 2>&1
echo "== fast sweep + fold"
timeout 300 $N -k regex:fast_sweep_kernel -s 3 -c 1 -o gpurun_out/ncu/fast_sweep_distmult python tools/run_sweep.py distmult 1024 14541 3 fast > gpurun_out/ncu/l7.log 2>&1
timeout 300 $N -k regex:fold_queries_kernel -s 3 -c 1 -o gpurun_out/ncu/fold_queries python tools/run_sweep.py distmult 1024 14541 3 fast > gpurun_out/ncu/l8.log 2>&1
echo "== train kernel B=1024 K=512 and B=64 K=512"
timeout 300 $NS -k regex:train_kernel -s 48 -c 1 -o gpurun_out/ncu/train_b1024_k512 python tools/run_train.py transe margin > gpurun_out/ncu/l9.log 2>&1
timeout 300 $N -k regex:train_kernel -s 25 -c 1 -o gpurun_out/ncu/train_b64_k512 python tools/run_train.py transe margin > gpurun_out/ncu/l10.log 2>&1
echo "== filter correction / store rows / sampler"
timeout 600 $N -k regex:filter_correct_indexed_kernel -s 3 -c 1 -o gpurun_out/ncu/filter_correct_indexed python tools/run_next_rows.py > gpurun_out/ncu/l11.log 2>&1
timeout 600 $N -k regex:store_rows_vec_kernel -s 3 -c 1 -o gpurun_out/ncu/store_rows_vec python tools/run_next_rows.py > gpurun_out/ncu/l12.log 2>&1
# the full reports are too large to travel together (64 MiB cap): export the raw / details pages here, keep only
# the report of the headline kernel
for f in gpurun_out/ncu/*.ncu-rep; do
  ncu -i $f --page raw --csv > ${f%.ncu-rep}.raw.csv 2>/dev/null
  ncu -i $f --page details --csv > ${f%.ncu-rep}.details.csv 2>/dev/null
done
ncu -i gpurun_out/ncu/sweep_transe_fb.ncu-rep --page source --csv > gpurun_out/ncu/sweep_transe_fb.source.csv 2>/dev/null
ncu -i gpurun_out/ncu/train_b1024_k512.ncu-rep --page source --csv > gpurun_out/ncu/train_b1024_k512.source.csv 2>/dev/null
find gpurun_out/ncu -name "*.ncu-rep" ! -name "sweep_transe_fb.ncu-rep" -delete
du -sh gpurun_out/ncu; ls -la gpurun_out/ncu | head -50
