#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1400 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_full.txt 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_full.txt | sed -E 's/ - .*//' | head -30
grep -E "^E  " gpurun_out/pytest_full.txt | sort | uniq -c | sort -rn | head -8 | cut -c1-300
python - <<'PY'
import time, torch, blp_b200
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(11)
n, r = 14541, 237
for d in (300, 768):
    ent = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=-1).to(dev)
    rel = ((torch.rand(r, d, generator=g) * 2 - 1) * 0.1).to(dev)
    for e in (64, 1024):
        tr = torch.stack([torch.randint(0, n, (e,), generator=g), torch.randint(0, n, (e,), generator=g), torch.randint(0, r, (e,), generator=g)], 1).to(dev)
        for _ in range(2): blp_b200.rank_sweep("transe", ent, rel, tr)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): blp_b200.rank_sweep("transe", ent, rel, tr)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(f"wide transe d={d} E={e}: {ms:.3f} ms/call, {2*e*n/ms/1e6:.1f} G scores/s, alu frac(2-op) {2*e*n*d*2/ms/1e9/36.7:.3f}")
PY
