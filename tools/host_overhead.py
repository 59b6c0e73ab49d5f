"""Host-side cost per call of the public API (GPU work is negligible at these sizes, so wall time ~= host time)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import blp_b200  # noqa: E402
from blp_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
n, d, b, k, T = 256, 128, 64, 64, 64
ent = torch.randn(n, d, generator=g).to(dev)
m = blp_b200.TransductiveLinkPrediction(d, "transe", "margin", n, 11, 0).to(dev)
x0 = torch.randn(b, 2, d, generator=g).to(dev)
rels = torch.randint(0, 11, (b, 1), generator=g).to(dev)
neg = blp_b200.get_negative_sampling_indices(b, k, device=dev)
rows = torch.stack([torch.randint(0, n, (T,), generator=g), torch.randint(0, n, (T,), generator=g),
                    torch.randint(0, 11, (T,), generator=g)], 1).to(dev)
step = blp_b200.GraphedLossStep(m, b, k)
step(x0, rels, neg)


def eager():
    x = x0.detach().requires_grad_(True)
    m.rel_emb.weight.grad = None
    loss = m.compute_loss(x, rels, neg)
    loss.backward()


def fwd_only():
    with torch.no_grad():
        m.compute_loss(x0, rels, neg)


cases = {
    "compute_loss fwd+bwd eager": eager,
    "compute_loss fwd only (no_grad)": fwd_only,
    "GraphedLossStep.replay": step.replay,
    "GraphedLossStep(x, rels, neg) (3 D2D copies + replay)": lambda: step(x0, rels, neg),
    "rank_sweep (T=64, N=256)": lambda: blp_b200.rank_sweep("transe", ent, m.rel_emb.weight, rows),
    "ops.rank_metrics": None,
    "RankSweepPlan (T=64, N=256)": None,
}
plan = blp_b200.RankSweepPlan("transe", ent, m.rel_emb.weight, T)
cases["RankSweepPlan (T=64, N=256)"] = lambda: plan(rows)
out = blp_b200.rank_sweep("transe", ent, m.rel_emb.weight, rows)
cases["ops.rank_metrics"] = lambda: ops.rank_metrics(out["gt"], out["ge"], (1, 3, 10))
for name, fn in cases.items():
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(300):
        fn()
    torch.cuda.synchronize()
    print(f"{name:55s} {(time.perf_counter() - t0) / 300 * 1e6:8.1f} us / call")
