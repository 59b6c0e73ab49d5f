#!/bin/bash
# N-GPU pass: NCCL sharded-sweep tests + bench under torchrun.  gpurun --gpus N --timeout 900 -- 'bash tools/gpu_multi.sh N'
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
echo "== pytest multi"; timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_multi.txt
echo "== bench x$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline 2> gpurun_out/bench_x$N.err | tee gpurun_out/bench_x$N.json | cut -c1-200; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_x$N.err | tail -5
echo "== bench reference arm x$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --impl reference --gpus $N --steps 2 --warmup 1 2> gpurun_out/bench_ref_x$N.err | tee gpurun_out/bench_ref_x$N.json | cut -c1-200
