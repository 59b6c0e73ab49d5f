#!/bin/bash
set -u
N=${1:-2}
mkdir -p gpurun_out
export BLP_BENCH_DEBUG=1
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null; nvidia-smi topo -m 2>/dev/null | head -8
echo "== pytest train"; timeout 900 python -m pytest tests/test_gpu_train.py -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.txt
echo "== bench x1"; timeout 600 python bench.py --steps 50 --warmup 5 --no-extra --no-cpu-baseline 2> gpurun_out/bench.err | cut -c1-200; grep rank gpurun_out/bench.err
echo "== bench x$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline --no-extra 2> gpurun_out/bench_x$N.err | tee gpurun_out/bench_x$N.json | cut -c1-200; grep rank gpurun_out/bench_x$N.err
echo "== bench x$N no flush"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline --no-extra --no-flush 2> gpurun_out/bench_x${N}_nf.err | cut -c1-200; grep rank gpurun_out/bench_x${N}_nf.err
