"""Where the time of one fused step goes: per-CTA globaltimer stamps written by the sweep kernel (blp_debug_timestamps).
    python tools/step_phases.py [model] [E] [N] [group_triples]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import blp_b200  # noqa: E402
from blp_b200 import _lib  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "transe"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 64
N = int(sys.argv[3]) if len(sys.argv) > 3 else 14541
group = int(sys.argv[4]) if len(sys.argv) > 4 else 0
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
ent = torch.randn(N, 128, generator=g)
if model == "transe":
    ent = torch.nn.functional.normalize(ent, dim=-1)
ent = ent.to(dev)
rel = ((torch.rand(237, 128, generator=g) * 2 - 1) * 0.128).to(dev)
triples = torch.stack([torch.randint(0, N, (E,), generator=g), torch.randint(0, N, (E,), generator=g),
                       torch.randint(0, 237, (E,), generator=g)], dim=1).contiguous().to(dev)
plan = blp_b200.RankSweepPlan(model, ent, rel, E, group_triples=group)
for _ in range(5):
    plan(triples)
torch.cuda.synchronize()
buf = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
_lib.lib().blp_debug_timestamps(ctypes.c_void_p(buf.data_ptr()))
for rep in range(3):
    buf.zero_()
    torch.cuda.synchronize()
    plan(triples)
    torch.cuda.synchronize()
    ts = buf.cpu().numpy().reshape(148, 16)
    ts = ts[ts[:, 0] > 0]
    t0 = ts[:, 0].min()
    names = ["start", "setup", "true", "fold", "tile0", "tiles_done", "ticket", "epilogue"]
    rel_us = (ts[:, :8].astype(np.float64) - t0) / 1e3
    rel_us[ts[:, :8] == 0] = np.nan
    print(f"rep {rep}: {len(ts)} CTAs, items/CTA min {ts[:, 9].min()} max {ts[:, 9].max()}, segments max {ts[:, 8].max()}")
    for k, nm in enumerate(names):
        col = rel_us[:, k]
        if np.all(np.isnan(col)):
            continue
        print(f"   {nm:11s} min {np.nanmin(col):7.2f}  median {np.nanmedian(col):7.2f}  max {np.nanmax(col):7.2f} us")
    dur = rel_us[:, 5] - rel_us[:, 4]
    per_item = dur / ts[:, 9]
    print(f"   tile loop per item: median {np.nanmedian(per_item):.2f} us, max {np.nanmax(per_item):.2f} us")
_lib.lib().blp_debug_timestamps(None)
