#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1400 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/pytest_full.txt 2>&1; tail -2 gpurun_out/pytest_full.txt
bash tools/gpu_bench.sh > gpurun_out/bench_run.txt 2>&1
grep -E "^(value|e2e|sharded)" gpurun_out/bench_run.txt | cut -c1-300
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print(d["wikidata5m_scale_sweep"]["table_pass_2"]); print(d["roofline"]["frac"], d["roofline"].get("wd_frac"))
PY
