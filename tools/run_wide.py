"""TransE sweep at a BOW encoder width through the public API (for ncu captures / timing): run_wide.py [d] [E] [N]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import blp_b200  # noqa: E402

d = int(sys.argv[1]) if len(sys.argv) > 1 else 768
E = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
N = int(sys.argv[3]) if len(sys.argv) > 3 else 14541
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(11)
ent = torch.nn.functional.normalize(torch.randn(N, d, generator=g), dim=-1).to(dev)
rel = ((torch.rand(237, d, generator=g) * 2 - 1) * 0.1).to(dev)
tr = torch.stack([torch.randint(0, N, (E,), generator=g), torch.randint(0, N, (E,), generator=g),
                  torch.randint(0, 237, (E,), generator=g)], 1).to(dev)
for _ in range(2):
    blp_b200.rank_sweep("transe", ent, rel, tr)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    out = blp_b200.rank_sweep("transe", ent, rel, tr)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
print(f"wide transe d={d} E={E} N={N}: {ms:.3f} ms/call, {2 * E * N / ms / 1e6:.1f} G scores/s, launches {out['launches']}")
