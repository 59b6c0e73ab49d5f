"""Run a few fused eval sweeps (for ncu captures and quick timing): python tools/run_sweep.py [model] [E] [N] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from blp_b200 import ops  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "transe"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
N = int(sys.argv[3]) if len(sys.argv) > 3 else 14541
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
ent = torch.randn(N, 128, generator=g)
if model == "transe":
    ent = torch.nn.functional.normalize(ent, dim=-1)
ent = ent.to(dev)
rel = ((torch.rand(237, 128, generator=g) * 2 - 1) * 0.128).to(dev)
h = ent[torch.randint(0, N, (E,), generator=g).to(dev)]
t = ent[torch.randint(0, N, (E,), generator=g).to(dev)]
r = rel[torch.randint(0, 237, (E,), generator=g).to(dev)]
for _ in range(3):
    ops.eval_rank(model, ent, h, t, r)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    out = ops.eval_rank(model, ent, h, t, r)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
print(f"{model} E={E} N={N}: {ms:.4f} ms per eval_rank call, {2 * E * N / ms / 1e6:.2f} G scores/s, gt[0]={int(out['gt'][0])}")
