#!/bin/bash
# gpurun --timeout 1500 -- 'bash tools/gpu_bench.sh'   : N=1 bench line (+ checksum), reference arm
set -u
mkdir -p gpurun_out
echo "== bench N=1"; timeout 900 python bench.py --steps 20 --warmup 5 --write-checksum 2> gpurun_out/bench.err > gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
except Exception as e:
    print("no bench line:", e); raise SystemExit
for k in ("value", "ms_per_step", "gpu_launches", "sharded_parity"):
    print(k, d.get(k))
print("e2e", d["e2e"]); print("clocks", d["clocks"]); print("cpu", d["cpu_baseline"])
print("roofline", json.dumps(d["roofline"], indent=1))
print("wd", json.dumps(d["wikidata5m_scale_sweep"], indent=1))
for k, v in (d.get("legs") or {}).items():
    print("leg", k, json.dumps(v)[:600])
PY
cp tests/golden/wd_sweep_checksum.json gpurun_out/ 2>/dev/null
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/bench_ref.err | tee gpurun_out/bench_reference.json | cut -c1-400
