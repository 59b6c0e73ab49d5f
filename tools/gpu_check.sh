#!/bin/bash
# One GPU-box pass: parity tests, bench lines, sweeps, next-row measurements, ncu launch list, ncu --set full captures.
# Run as: gpurun --timeout 1800 -- 'bash tools/gpu_check.sh'   (outputs under gpurun_out/)
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.txt
echo "== probe"; timeout 300 python tools/probe_fp32.py 2>&1 | tee gpurun_out/probe_fp32.txt
echo "== bench exact transe"; timeout 600 python bench.py --steps 50 --warmup 5 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-300
echo "== bench fast distmult"; timeout 600 python bench.py --steps 50 --warmup 5 --model distmult --mode fast --no-extra 2> gpurun_out/bench_fast.err | tee gpurun_out/bench_fast_distmult.json | cut -c1-300
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2> gpurun_out/bench_ref.err | tee gpurun_out/bench_reference.json | cut -c1-300
echo "== sweeps"
rm -f gpurun_out/sweeps.txt
for spec in "transe 64 14541 30" "distmult 64 14541 30" "transe 1024 14541 20" "distmult 1024 14541 10" "complex 1024 40943 5" "simple 1024 14541 10" "transe 2 4800000 10" "transe 64 4800000 3" \
            "distmult 64 14541 30 fast" "distmult 1024 14541 20 fast" "complex 1024 40943 10 fast" "simple 1024 14541 20 fast" "distmult 8192 14541 10 fast" "distmult 64 4800000 3 fast"; do
  timeout 300 python tools/run_sweep.py $spec 2>&1 | tail -1 | tee -a gpurun_out/sweeps.txt
done
echo "== train kernel"; timeout 300 python tools/run_train.py transe margin 2>&1 | tee gpurun_out/train_kernel.txt
echo "== next rows"; timeout 600 python tools/run_next_rows.py 2>&1 | tee gpurun_out/next_rows.txt
echo "== host overhead"; timeout 300 python tools/host_overhead.py 2>&1 | grep "us / call" | tee gpurun_out/host_overhead.txt
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_under_ncu.log 2>&1
echo "== ncu full: sweep transe FB (from the bench command itself)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 6 -c 1 -f -o gpurun_out/sweep_transe_fb \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu_fb.log 2>&1
echo "== ncu full: sweep transe WD (HBM-bound, eval batch 2)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 3 -c 1 -f -o gpurun_out/sweep_transe_wd \
  python tools/run_sweep.py transe 2 4800000 2 > gpurun_out/ncu_wd.log 2>&1
echo "== ncu full: fast sweep distmult FB (tcgen05)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fast_sweep_kernel -s 3 -c 1 -f -o gpurun_out/sweep_fast_distmult_fb \
  python tools/run_sweep.py distmult 1024 14541 2 fast > gpurun_out/ncu_fast.log 2>&1
ls -la gpurun_out | head -50
