#!/bin/bash
set -u
python - <<'PY'
import torch, blp_b200
from blp_b200 import ops
for rows in (2048, 128, 16384, 262144):
    for iters in (64, 256):
        g, t = ops.atomic_probe("cuda:0", rows=rows, iters=iters)
        print(f"atomic probe rows={rows} iters={iters}: {g:.0f} GB/s")
PY
