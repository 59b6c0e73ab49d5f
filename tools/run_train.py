"""Time the fused compute_loss kernel (forward + backward) at a few batch shapes:
    python tools/run_train.py [model] [loss]
Reports us per launch (CUDA events, kernel bracketed by blp_profile_events) against the HBM / FP32 floors."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import blp_b200  # noqa: E402
from blp_b200 import _lib, ops  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "transe"
loss = sys.argv[2] if len(sys.argv) > 2 else "margin"
dev = torch.device("cuda", 0)
lib = _lib.lib()
d, n_rel = 128, 237
for b, k in ((64, 64), (64, 512), (1024, 512), (8192, 64)):
    g = torch.Generator().manual_seed(0)
    ent = torch.randn(b, 2, d, generator=g)
    if model == "transe":
        ent = torch.nn.functional.normalize(ent, dim=-1)
    ent = ent.to(dev)
    rel = ((torch.rand(n_rel, d, generator=g) * 2 - 1) * 0.128).to(dev)
    rels = torch.randint(0, n_rel, (b, 1), generator=g).to(dev)
    neg = blp_b200.get_negative_sampling_indices(b, k, device=dev, seed=1)
    reps = 20
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a_, b_ in evs:
        a_.record(); b_.record()
    for _ in range(3):
        ops.train_loss(model, loss, ent, rel, rels, neg, want_grad=True)
    torch.cuda.synchronize()
    for a_, b_ in evs:
        lib.blp_profile_events(2, ctypes.c_void_p(a_.cuda_event), ctypes.c_void_p(b_.cuda_event))
        ops.train_loss(model, loss, ent, rel, rels, neg, want_grad=True)
    lib.blp_profile_events(0, None, None)
    torch.cuda.synchronize()
    us = sorted(a_.elapsed_time(b_) * 1e3 for a_, b_ in evs)[reps // 2]
    alg = 2 * b * d * 4 + b * d * 4 + b * 8 + b * k * 16 + 3 * b * d * 4
    lane = b * (k + 1) * d * {"transe": 6, "distmult": 8, "complex": 30, "simple": 14}[model]
    print(f"{model}/{loss} B={b} K={k}: {us:.1f} us per fused fwd+bwd kernel, {b * (k + 1) / us:.1f} M triples/s, "
          f"HBM floor {alg / 6550.7e3:.2f} us, FP32 floor ~{lane / 36.5e6:.2f} us")
