#!/bin/bash
# N-GPU pass: gpurun --gpus N --timeout 1500 -- 'bash tools/gpu_multi.sh N'
set -u
N=${1:-2}
mkdir -p gpurun_out
echo "== NCCL tests"; timeout 600 python -m pytest tests/test_gpu_multi.py -q --tb=short 2>&1 | tail -5 | tee gpurun_out/pytest_multi.txt
echo "== bench x$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/bench_x$N.err > gpurun_out/bench_x$N.json
tail -3 gpurun_out/bench_x$N.err
python - "$N" <<'PY'
import json, sys
n = sys.argv[1]
d = json.loads(open(f"gpurun_out/bench_x{n}.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", d["sharded_parity"])
print(json.dumps(d["wikidata5m_scale_sweep"], indent=1))
PY
