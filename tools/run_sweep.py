"""Run a few fused eval sweeps (for ncu captures and quick timing):
    python tools/run_sweep.py [model] [E] [N] [reps] [exact|fast|fast_exact]
E test triples (2E queries) are ranked against N candidates per call through blp_rank_sweep[_fast]."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from blp_b200 import ops  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "transe"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
N = int(sys.argv[3]) if len(sys.argv) > 3 else 14541
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
mode = sys.argv[5] if len(sys.argv) > 5 else "exact"
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
ent = torch.randn(N, 128, generator=g)
if model == "transe":
    ent = torch.nn.functional.normalize(ent, dim=-1)
ent = ent.to(dev)
rel = ((torch.rand(237, 128, generator=g) * 2 - 1) * 0.128).to(dev)
triples = torch.stack([torch.randint(0, N, (E,), generator=g), torch.randint(0, N, (E,), generator=g),
                       torch.randint(0, 237, (E,), generator=g)], dim=1).to(dev)
if os.environ.get("SORT_REL"):
    triples = triples[torch.argsort(triples[:, 2], stable=True)].contiguous()
out = {k: torch.empty((2, E), dtype=torch.int32, device=dev) for k in ("gt", "ge")}
out["true_score"] = torch.empty((2, E), dtype=torch.float32, device=dev)
ws = ops.fast_table(ent) if mode.startswith("fast") else None
rws = ops.refine_workspace(dev, max(1 << 19, 128 * E)) if mode == "fast_exact" else None


def call():
    return ops.rank_sweep_chunk(model, ent, rel, triples, out, 0, E, fast_table_ws=ws, refine_ws=rws)


for _ in range(3):
    call()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    call()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
alg = N * 128 * 4 + E * 3 * 128 * 4 + 2 * E * 12
print(f"{model} {mode} E={E} N={N}: {ms:.4f} ms per rank_sweep call, {2 * E * N / ms / 1e6:.2f} G scores/s, "
      f"{alg / ms / 1e6:.1f} GB/s algorithmic, gt[0]={int(out['gt'][0, 0])}"
      + (f", refine entries {int(rws[:8].view(torch.int32)[0])} overflow {int(rws[:8].view(torch.int32)[1])}" if rws is not None else ""))
