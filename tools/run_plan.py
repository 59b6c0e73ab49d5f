"""Time RankSweepPlan calls (true scores + sweep + metrics per call), for ncu launch lists of a whole call:
    python tools/run_plan.py [model] [E] [N] [reps] [exact|fast|fast_exact]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import blp_b200  # noqa: E402
from blp_b200 import ops  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "distmult"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 20480
N = int(sys.argv[3]) if len(sys.argv) > 3 else 14541
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
mode = sys.argv[5] if len(sys.argv) > 5 else "fast"
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
ent = torch.randn(N, 128, generator=g)
if model == "transe":
    ent = torch.nn.functional.normalize(ent, dim=-1)
ent = ent.to(dev)
rel = ((torch.rand(237, 128, generator=g) * 2 - 1) * 0.128).to(dev)
triples = torch.stack([torch.randint(0, N, (E,), generator=g), torch.randint(0, N, (E,), generator=g),
                       torch.randint(0, 237, (E,), generator=g)], dim=1).to(dev)
plan = blp_b200.RankSweepPlan(model, ent, rel, E, mode=mode, fast_table=ops.fast_table(ent) if mode.startswith("fast") else None)
for _ in range(3):
    out = plan(triples)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    out = plan(triples)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
print(f"{model} {mode} plan E={E} N={N}: {ms * 1e3:.1f} us per call, launches/call {out['launches']}, "
      f"{2 * E * N / ms / 1e6:.2f} G scores/s, sums {[round(float(x), 6) for x in out['sums'][:3]]}")
