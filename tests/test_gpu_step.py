"""The fused one-launch step (blp_rank_step / blp_plan_run / blp_rank_queries) vs the oracle: true scores computed
inside the sweep kernel, zero-invariant workspace, last-CTA epilogue (counters + get_metrics + sums), forced table-pass
sizes (the reference's eval_batch_size), independent head / tail query sets.  Bit-exact, through the C ABI."""
import numpy as np
import pytest
import torch

from oracle import c_oracle

import blp_b200
from blp_b200 import ops
from test_gpu_eval import MODELS, make_inputs

pytestmark = pytest.mark.gpu


def _oracle(model, ent, rel, heads, tails, rels):
    co = c_oracle.eval_rank(model, ent.numpy(), ent[heads].numpy(), ent[tails].numpy(), rel[rels].numpy(),
                            heads.numpy(), tails.numpy())
    recip, hits = c_oracle.metrics_from_counts(co["gt"], co["ge"], [1, 3, 10])
    return co, recip, hits


def _step(model, ent, rel, triples, dev, group_triples=0, metrics=True):
    T = triples.shape[0]
    out = {k: torch.full((2, T), -7, dtype=torch.int32, device=dev) for k in ("gt", "ge")}
    out["true_score"] = torch.empty((2, T), dtype=torch.float32, device=dev)
    m = ops.alloc_metrics(dev, 2 * T, 3) if metrics else None
    launches = ops.rank_step(model, ent, rel, triples, out, group_triples=group_triples, k_values=(1, 3, 10), metrics=m)
    return out, m, launches


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("n,b", [(1, 1), (31, 2), (257, 3), (129, 65), (1000, 7), (777, 12), (300, 33), (14541, 64), (5000, 300)])
def test_rank_step_vs_oracle(model, n, b, cuda_device):
    ent, rel, heads, tails, rels = make_inputs(model, n, 128, b, seed=n * 3 + b)
    if n > 40:
        ent[n - 1] = ent[3]            # exact ties
        heads[0] = 3
    co, recip, hits = _oracle(model, ent, rel, heads, tails, rels)
    triples = torch.stack([heads, tails, rels], 1).to(cuda_device)
    e, r = ent.to(cuda_device), rel.to(cuda_device)
    for rep in range(2):               # second call: the workspace was left zeroed by the first
        out, m, launches = _step(model, e, r, triples, cuda_device)
        assert launches == 1
        assert np.array_equal(out["true_score"].reshape(-1).cpu().numpy(), co["true_score"])
        assert np.array_equal(out["gt"].reshape(-1).cpu().numpy(), co["gt"])
        assert np.array_equal(out["ge"].reshape(-1).cpu().numpy(), co["ge"])
        assert np.array_equal(m["recip"].cpu().numpy().reshape(-1), recip.reshape(-1))
        assert np.array_equal(m["hits"].cpu().numpy().astype(bool), hits.astype(bool))
        sums = m["sums"].cpu().numpy()
        assert abs(sums[0] - recip.astype(np.float64).sum()) < 1e-9
        assert np.array_equal(sums[1:], hits.sum(0).astype(np.float64))


@pytest.mark.parametrize("model", ("transe", "complex"))
@pytest.mark.parametrize("group", (2, 4, 8, 16, 32))
def test_rank_step_forced_table_pass_size(model, group, cuda_device):
    """group_triples = the reference's eval_batch_size: every pass over the table serves `group` triples; ranks do not
    depend on it."""
    n, b = 3000, 37
    ent, rel, heads, tails, rels = make_inputs(model, n, 128, b, seed=group)
    co, _, _ = _oracle(model, ent, rel, heads, tails, rels)
    triples = torch.stack([heads, tails, rels], 1).to(cuda_device)
    out, _, launches = _step(model, ent.to(cuda_device), rel.to(cuda_device), triples, cuda_device, group_triples=group)
    assert launches == 1
    assert np.array_equal(out["gt"].reshape(-1).cpu().numpy(), co["gt"])
    assert np.array_equal(out["ge"].reshape(-1).cpu().numpy(), co["ge"])


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("group", (2, 4))
def test_tile_split_pass_sizes_on_many_tiles(model, group, cuda_device):
    """Pass sizes 2 / 4 run the tile-split register tiles (both slots of a CTA rank the same triples on alternating candidate
    tiles, one `full` barrier per (slot, buffer)): many tiles per CTA, a ragged last tile, an odd number of triples, three
    repetitions (the first version raced on its barrier phases once in a few hundred launches)."""
    n, b = 60_001, 7
    ent, rel, heads, tails, rels = make_inputs(model, n, 128, b, seed=40 + group)
    co, _, _ = _oracle(model, ent, rel, heads, tails, rels)
    triples = torch.stack([heads, tails, rels], 1).to(cuda_device)
    e, r = ent.to(cuda_device), rel.to(cuda_device)
    for _ in range(3):
        out, _, launches = _step(model, e, r, triples, cuda_device, group_triples=group)
        assert launches == 1
        assert np.array_equal(out["true_score"].reshape(-1).cpu().numpy(), co["true_score"])
        assert np.array_equal(out["gt"].reshape(-1).cpu().numpy(), co["gt"])
        assert np.array_equal(out["ge"].reshape(-1).cpu().numpy(), co["ge"])


def test_rank_step_long_output_range_uses_separate_metrics(cuda_device):
    """2T > 16384: counters zeroed by memset, metrics by their own launch; same results."""
    n, b = 300, 9000
    ent, rel, heads, tails, rels = make_inputs("distmult", n, 128, b, seed=1)
    co, recip, hits = _oracle("distmult", ent, rel, heads, tails, rels)
    triples = torch.stack([heads, tails, rels], 1).to(cuda_device)
    out, m, launches = _step("distmult", ent.to(cuda_device), rel.to(cuda_device), triples, cuda_device)
    assert launches == 2
    assert np.array_equal(out["gt"].reshape(-1).cpu().numpy(), co["gt"])
    assert np.array_equal(out["ge"].reshape(-1).cpu().numpy(), co["ge"])
    assert np.array_equal(m["recip"].cpu().numpy().reshape(-1), recip.reshape(-1))


def test_rank_step_out_of_range_index_is_flagged(cuda_device):
    ent, rel, heads, tails, rels = make_inputs("transe", 500, 128, 5, seed=2)
    triples = torch.stack([heads, tails, rels], 1)
    triples[2, 0] = 500                       # not a table row (train.py:137-138 asserts this never happens)
    out, _, _ = _step("transe", ent.to(cuda_device), rel.to(cuda_device), triples.to(cuda_device), cuda_device)
    assert torch.isnan(out["true_score"][0, 2]) and int(out["gt"][0, 2]) == 0 and int(out["ge"][0, 2]) == 0
    assert not torch.isnan(out["true_score"][0, 1])


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("n_hq,n_tq", [(20, 20), (64, 64), (7, 0), (0, 9), (5, 12)])
def test_rank_queries_independent_sets(model, n_hq, n_tq, cuda_device):
    """Head-prediction and tail-prediction queries that do NOT come from the same triples (the score-matrix path)."""
    n = 2000
    g = torch.Generator().manual_seed(n_hq * 31 + n_tq)
    ent, rel, _, _, _ = make_inputs(model, n, 128, 1, seed=5)
    hq_true, hq_tail, hq_rel = (torch.randint(0, m_, (n_hq,), generator=g) for m_ in (n, n, 11))
    tq_true, tq_head, tq_rel = (torch.randint(0, m_, (n_tq,), generator=g) for m_ in (n, n, 11))
    dev = cuda_device
    e = ent.to(dev)
    res = ops.rank_queries(model, e,
                           (ent[hq_tail].to(dev), rel[hq_rel].to(dev), hq_true.to(dev)) if n_hq else None,
                           (ent[tq_head].to(dev), rel[tq_rel].to(dev), tq_true.to(dev)) if n_tq else None,
                           k_values=(1, 3, 10))
    # oracle: two independent batches, take the head half of the first and the tail half of the second
    want_gt, want_ge, want_ts = [], [], []
    if n_hq:
        co = c_oracle.eval_rank(model, ent.numpy(), ent[hq_true].numpy(), ent[hq_tail].numpy(), rel[hq_rel].numpy(),
                                hq_true.numpy(), hq_tail.numpy())
        want_gt.append(co["gt"][:n_hq]); want_ge.append(co["ge"][:n_hq]); want_ts.append(co["true_score"][:n_hq])
    if n_tq:
        co = c_oracle.eval_rank(model, ent.numpy(), ent[tq_head].numpy(), ent[tq_true].numpy(), rel[tq_rel].numpy(),
                                tq_head.numpy(), tq_true.numpy())
        want_gt.append(co["gt"][n_tq:]); want_ge.append(co["ge"][n_tq:]); want_ts.append(co["true_score"][n_tq:])
    gt, ge, ts = (np.concatenate(x) for x in (want_gt, want_ge, want_ts))
    assert np.array_equal(res["true_score"].cpu().numpy(), ts)
    assert np.array_equal(res["gt"].cpu().numpy(), gt) and np.array_equal(res["ge"].cpu().numpy(), ge)
    recip, hits = c_oracle.metrics_from_counts(gt, ge, [1, 3, 10])
    assert np.array_equal(res["recip"].cpu().numpy().reshape(-1), recip.reshape(-1))
    assert np.array_equal(res["hits"].cpu().numpy(), hits.astype(bool))
    assert res["launches"] == (1 if n_hq == n_tq else (1 if 0 in (n_hq, n_tq) else 2) + 1)


@pytest.mark.parametrize("model", ("transe", "simple"))
def test_plan_one_launch_per_batch(model, cuda_device):
    """RankSweepPlan = blp_plan_run: one ctypes call, one launch per batch of 64 (the reference's eval batch)."""
    n, T = 14541, 64
    ent, rel, heads, tails, rels = make_inputs(model, n, 128, 3 * T, seed=8, n_rel=237)
    e, r = ent.to(cuda_device), rel.to(cuda_device)
    plan = blp_b200.RankSweepPlan(model, e, r, T)
    for c in range(3):
        sl = slice(c * T, (c + 1) * T)
        triples = torch.stack([heads[sl], tails[sl], rels[sl]], 1).to(cuda_device)
        out = plan(triples)
        assert out["launches"] == 1
        co, recip, hits = _oracle(model, ent, rel, heads[sl], tails[sl], rels[sl])
        assert np.array_equal(out["gt"].reshape(-1).cpu().numpy(), co["gt"])
        assert np.array_equal(out["ge"].reshape(-1).cpu().numpy(), co["ge"])
        assert np.array_equal(out["true_score"].reshape(-1).cpu().numpy(), co["true_score"])
        assert np.array_equal(out["recip"].cpu().numpy().reshape(-1), recip.reshape(-1))
        assert abs(float(out["sums"][0]) - recip.astype(np.float64).sum()) < 1e-9


@pytest.mark.parametrize("model", ("transe", "distmult"))
def test_plan_overlapping_calls_equal_serial_calls(model, cuda_device):
    """RankSweepPlan(overlap_calls=True): batch n + 1 is launched with programmatic stream serialization and starts while
    batch n is still finishing; it must not touch the shared workspace / outputs before batch n is complete.  Every
    batch of a back-to-back sequence must give the serial plan's numbers (checked through asynchronous copies between
    the calls, and -- pure kernel-after-kernel -- through the state the sequence leaves behind)."""
    n, e, n_batches = 14541, 64, 6
    ent, rel, heads, tails, rels = make_inputs(model, n, 128, e * n_batches, seed=91, n_rel=37)
    dev = cuda_device
    en, rl = ent.to(dev), rel.to(dev)
    triples = torch.stack([heads, tails, rels], 1).to(dev)
    batches = [triples[i * e:(i + 1) * e] for i in range(n_batches)]          # slices of a resident tensor
    serial = blp_b200.RankSweepPlan(model, en, rl, e)
    want = []
    for bt in batches:
        o = serial(bt)
        want.append({k: o[k].clone() for k in ("gt", "ge", "true_score", "sums")})
    torch.cuda.synchronize()
    plan = blp_b200.RankSweepPlan(model, en, rl, e, overlap_calls=True)
    got = []
    for bt in batches:                                 # copies between the calls
        o = plan(bt)
        got.append({k: o[k].clone() for k in ("gt", "ge", "true_score", "sums")})
    torch.cuda.synchronize()
    for w, g in zip(want, got):
        for k in w:
            assert torch.equal(w[k], g[k]), k
    for rep in range(40):                              # kernel directly after kernel
        o = plan(batches[rep % n_batches])
    o = plan(batches[2])
    torch.cuda.synchronize()
    for k in want[2]:
        assert torch.equal(want[2][k], o[k]), k
