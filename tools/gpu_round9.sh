#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/run_train.py transe margin 2>&1 | tee gpurun_out/train_kernel.txt
timeout 300 python tools/run_train.py complex nll 2>&1 | tee -a gpurun_out/train_kernel.txt
echo "== bench exact"; timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-200; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
r=d["roofline"]; print("ms/step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "kern_ms", r["kernel_ms"], "alu frac", r["alu"]["frac"])
print(d["wikidata5m_scale_sweep"]["eval_batch_2"])
PY
timeout 120 python tools/run_sweep.py transe 2 4800000 10
