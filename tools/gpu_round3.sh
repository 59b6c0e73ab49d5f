#!/bin/bash
set -u
N=${1:-1}
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.txt
echo "== host overhead"; timeout 300 python tools/host_overhead.py 2>&1 | tee gpurun_out/host_overhead.txt
echo "== bench exact"; timeout 600 python bench.py --steps 50 --warmup 5 --no-extra 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-400; tail -3 gpurun_out/bench.err
echo "== bench fast distmult"; timeout 600 python bench.py --steps 50 --warmup 5 --model distmult --mode fast --no-extra --no-cpu-baseline 2> gpurun_out/bench_fast.err | tee gpurun_out/bench_fast_distmult.json | cut -c1-400; tail -3 gpurun_out/bench_fast.err
if [ "$N" -gt 1 ]; then
echo "== bench x$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline --no-extra 2> gpurun_out/bench_x$N.err | tee gpurun_out/bench_x$N.json | cut -c1-400; tail -5 gpurun_out/bench_x$N.err
fi
