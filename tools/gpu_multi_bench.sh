#!/bin/bash
# N-GPU bench only: gpurun --gpus N --timeout 900 -- 'bash tools/gpu_multi_bench.sh N'
set -u
N=${1:-8}
mkdir -p gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/bench_x$N.err > gpurun_out/bench_x$N.json
tail -3 gpurun_out/bench_x$N.err
python - "$N" <<'PY'
import json, sys
n = sys.argv[1]
d = json.loads(open(f"gpurun_out/bench_x{n}.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", d["sharded_parity"], "clocks", d["clocks"])
print(json.dumps(d["wikidata5m_scale_sweep"]["table_pass_2"]))
PY
