#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fast.py -q --tb=short -x 2>&1 | tail -5
for m in fast fast_exact; do
  timeout 120 python tools/run_sweep.py distmult 16384 14541 5 $m 2>&1 | tail -1
  timeout 120 python tools/run_sweep.py complex 3136 40943 5 $m 2>&1 | tail -1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fast_exact.csv \
  python tools/run_sweep.py distmult 16384 14541 2 fast_exact > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_fast_exact.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
for r in rows[-4:]:
    print(r[ki][:60], r[vi], r[ui])
PY
