"""Measure the FP32 pipe rates the exact-order sweeps are bound by (blp_pipe_probe variants)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from blp_b200 import ops  # noqa: E402

NAMES = ["FADD", "FADD+|x| (TransE tail step)", "add.f32x2", "f32x2 TransE step (2 packed adds + LOP)", "FMUL+FADD (DistMult tail step)", "f32x2 DistMult tail step (FFMA2 + FADD2)"]
dev = torch.device("cuda", 0)
for v, name in enumerate(NAMES):
    best = 0.0
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sink = None
        a.record()
        lane_ops, sink = ops.pipe_probe(v, dev, n_threads=148 * 8 * 256, iters=16384)
        b.record()
        torch.cuda.synchronize()
        best = max(best, lane_ops / (a.elapsed_time(b) * 1e-3) / 1e12)
    print(f"variant {v} {name:45s} {best:7.2f} T lane-ops/s")
