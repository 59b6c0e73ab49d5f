"""Measure the rows either side of the hot path on the GPU box, next to CPU restatements of the
reference's way of doing them (oracle/torch_port.py; test infrastructure, used here only as the timed baseline):
    f1  filter index build + filtered-rank correction     vs  per-batch dense masks (utils.py:46-83) + H2D + re-rank
    f2  MRR breakdowns on the device                      vs  per-triple Python loops (utils.py:114-168)
    f3  negative sampler                                  vs  CPU multinomial (data.py:35-81) + H2D
    f4  normalise + scatter into a row shard              vs  F.normalize + slice assignment (models.py:40, train.py:112)
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import blp_b200  # noqa: E402
from blp_b200 import ops  # noqa: E402
from oracle import torch_port  # noqa: E402

dev = torch.device("cuda", 0)
PEAK = 6550.7


def gpu_ms(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def cpu_ms(fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


# ---------------------------------------------------------------- f1: filtered ranks (FB15k-237-shaped) ----
n, n_rel, n_edges, T = 14541, 237, 310116, 1024
rng = np.random.default_rng(0)
# N-to-N structure: heads drawn from a small hub set so (head, rel) keys have tens of known tails
hubs = rng.integers(0, n, 2000)
edges = np.stack([hubs[rng.integers(0, 2000, n_edges)], rng.integers(0, n, n_edges), rng.integers(0, 20, n_edges)], 1)
test = edges[rng.integers(0, n_edges, T)].copy()
g = torch.Generator().manual_seed(0)
ent = torch.nn.functional.normalize(torch.randn(n, 128, generator=g), dim=-1).to(dev)
rel = ((torch.rand(n_rel, 128, generator=g) * 2 - 1) * 0.128).to(dev)
rows = torch.from_numpy(test).to(dev)
ms_build = gpu_ms(lambda: blp_b200.DeviceFilterIndex(edges, None, n, n_rel, dev), reps=5, warm=1)
didx = blp_b200.DeviceFilterIndex(edges, None, n, n_rel, dev)
plan_raw = blp_b200.RankSweepPlan("transe", ent, rel, T)
plan_f = blp_b200.RankSweepPlan("transe", ent, rel, T, filter_index=didx)
ms_raw, ms_filt = gpu_ms(lambda: plan_raw(rows)), gpu_ms(lambda: plan_f(rows))
out = plan_f(rows)
removed = float((out["ge"] - out["ge_f"]).float().mean())
out_edges, in_edges = {}, {}
for h, t, r in edges.tolist():
    out_edges.setdefault(h, []).append((h, t, r))
    in_edges.setdefault(t, []).append((h, t, r))
ent2idx = torch.arange(n)
batch = torch.from_numpy(test[:64])


def ref_filter_batch():
    hm, tm = torch_port.triple_filter_masks(batch, out_edges, in_edges, n, ent2idx)
    return torch.cat((hm, tm)).to(dev)


ms_ref = cpu_ms(ref_filter_batch)
print(f"f1 filter index build ({n_edges} edges, once per evaluation): {ms_build:.3f} ms")
print(f"f1 sweep of {T} triples raw {ms_raw:.3f} ms, raw + filtered {ms_filt:.3f} ms (correction {1e3 * (ms_filt - ms_raw):.1f} us, "
      f"{removed:.1f} filtered candidates per query); reference-style mask build + H2D: {ms_ref:.1f} ms per batch of 64 "
      f"= {ms_ref * T / 64:.0f} ms per {T} triples")

# ---------------------------------------------------------------- f2: breakdowns ----
T2 = 20480
trip_ids = torch.stack([torch.randint(0, n, (T2,), generator=g), torch.randint(0, n, (T2,), generator=g),
                        torch.randint(0, n_rel, (T2,), generator=g)], 1)
recip = torch.rand(2 * T2, generator=g)
new = set(torch.randint(0, n, (3000,), generator=g).tolist())
is_new = torch.zeros(n, dtype=torch.uint8)
is_new[list(new)] = 1
cats = torch.randint(0, 4, (n_rel,), generator=g)
d_recip, d_trip, d_new, d_cats = recip.to(dev), trip_ids.to(dev), is_new.to(dev), cats.to(dev)
ms_bd = gpu_ms(lambda: ops.mrr_breakdown(d_recip, d_trip, d_new, d_cats))
ms_bd_ref = cpu_ms(lambda: torch_port.mrr_by_new_position(trip_ids[:2048], torch.cat([recip[:2048], recip[T2:T2 + 2048]]), new), reps=1)
print(f"f2 MRR breakdowns over {T2} triples: {1e3 * ms_bd:.1f} us on the device; reference-style Python loop: "
      f"{ms_bd_ref * T2 / 2048:.0f} ms (by position only, extrapolated from 2048 triples)")

# ---------------------------------------------------------------- f3: negative sampler ----
for b, k in ((64, 512), (1024, 512)):
    ms = gpu_ms(lambda: blp_b200.get_negative_sampling_indices(b, k, device=dev, seed=1, offset=3))
    ms_ref = cpu_ms(lambda: torch_port.sample_negative_indices(b, k).contiguous().to(dev))
    nbytes = b * k * 16
    print(f"f3 sampler B={b} K={k}: {1e3 * ms:.1f} us ({nbytes / ms / 1e6:.0f} GB/s written, {100 * nbytes / ms / 1e6 / PEAK:.1f}% of HBM peak); "
          f"reference-style CPU multinomial + H2D: {ms_ref:.2f} ms")

# ---------------------------------------------------------------- f4: normalise + scatter ----
for m in (14541, 600000, 4800000):
    raw = torch.randn(m, 128, device=dev)
    shard = torch.empty_like(raw)
    ms = gpu_ms(lambda: blp_b200.store_rows(shard, raw, normalize=True), reps=10)
    nbytes = 2 * m * 512
    line = f"f4 normalise + store {m} rows: {ms:.3f} ms ({nbytes / ms / 1e6:.0f} GB/s, {100 * nbytes / ms / 1e6 / PEAK:.1f}% of HBM peak)"
    if m <= 600000:
        rc = raw.cpu()
        dst = torch.empty_like(rc)

        def ref_norm():
            dst[:] = torch_port.normalize_rows(rc)
        line += f"; F.normalize + slice assignment on the CPU: {cpu_ms(ref_norm):.1f} ms"
    print(line)
    del raw, shard
