"""Per-call time of the fused one-launch step through RankSweepPlan (blp_plan_run), back to back without syncs:
    python tools/run_step.py [model] [E] [N] [reps] [group_triples]
Prints GPU time per call (CUDA events over the whole loop) and host time per call (enqueue only)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import blp_b200  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "transe"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 64
N = int(sys.argv[3]) if len(sys.argv) > 3 else 14541
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 200
group = int(sys.argv[5]) if len(sys.argv) > 5 else 0
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
ent = torch.randn(N, 128, generator=g)
if model == "transe":
    ent = torch.nn.functional.normalize(ent, dim=-1)
ent = ent.to(dev)
rel = ((torch.rand(237, 128, generator=g) * 2 - 1) * 0.128).to(dev)
triples = torch.stack([torch.randint(0, N, (E,), generator=g), torch.randint(0, N, (E,), generator=g),
                       torch.randint(0, 237, (E,), generator=g)], dim=1)
if os.environ.get("SORT_REL"):
    triples = triples[torch.argsort(triples[:, 2], stable=True)]
triples = triples.contiguous().to(dev)
plan = blp_b200.RankSweepPlan(model, ent, rel, E, group_triples=group, overlap_calls=bool(os.environ.get("OVERLAP")))
for _ in range(5):
    out = plan(triples)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
a.record()
for _ in range(reps):
    out = plan(triples)
b.record()
host = (time.perf_counter() - t0) / reps
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
# one call timed alone (launch to completion), median of 20
lat = []
for _ in range(20):
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    out = plan(triples)
    torch.cuda.synchronize()
    lat.append(time.perf_counter() - t1)
lat.sort()
print(f"{model} step E={E} N={N} group={group}: {ms * 1e3:.1f} us per call back to back (host enqueue {host * 1e6:.1f} us), "
      f"{lat[10] * 1e6:.1f} us alone incl. sync, launches/call {out['launches']}, {2 * E * N / ms / 1e6:.2f} G scores/s, "
      f"mrr {float(out['sums'][0]) / (2 * E):.6f}")
