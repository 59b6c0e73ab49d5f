#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== bench exact"; timeout 600 python bench.py --steps 50 --warmup 5 --no-extra --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-200; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
r=d["roofline"]; print("ms/step", d["ms_per_step"], "kern_ms", r["kernel_ms"], "alu frac", r["alu"]["frac"])
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 6 -c 1 -f -o gpurun_out/sweep_transe_fb_sorted python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > /dev/null 2>&1
