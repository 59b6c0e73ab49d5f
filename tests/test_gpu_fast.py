"""Tensor-core (tcgen05, split-FP16) mode of the eval sweep vs the oracle: tolerance-classified parity.

The fast mode folds the query side (r*t, h*r, ...) and sums in tensor-core order, so it cannot carry the
reference's fp32 roundings (models.py:227 association).  Contract (north_star "scores within
1e-5 relative"): |s_fast - s_ref| <= 1e-5 * sum_d |term_d|, and a rank may differ from the exact one only by
candidates whose reference score lies within that tolerance of the true score."""
import numpy as np
import pytest
import torch

from oracle import c_oracle, np_oracle

import blp_b200
from blp_b200 import ops
from test_gpu_eval import make_inputs

pytestmark = pytest.mark.gpu
FAST_MODELS = ("distmult", "complex", "simple")
TOL = 1e-5


def _run_fast(model, ent, rel, heads, tails, rels, dev, want_scores=True, ent_offset=0, shard=None):
    e = ent.to(dev)
    table = e if shard is None else e[shard[0]:shard[1]].contiguous()
    T = heads.numel()
    triples = torch.stack([heads, tails, rels], dim=1).to(dev)
    out = {k: torch.empty((2, T), dtype=torch.int32, device=dev) for k in ("gt", "ge")}
    out["true_score"] = torch.empty((2, T), dtype=torch.float32, device=dev)
    scores = torch.full((2 * T, table.shape[0]), float("nan"), device=dev) if want_scores else None
    kw = {}
    if shard is not None:
        kw = {"h_rows": e[triples[:, 0]], "t_rows": e[triples[:, 1]]}
    ops.rank_sweep_chunk(model, table, rel.to(dev), triples, out, 0, T, ent_offset=ent_offset,
                         fast_table_ws=ops.fast_table(table), scores_out=scores, **kw)
    torch.cuda.synchronize()
    return out, scores


@pytest.mark.parametrize("model", FAST_MODELS)
@pytest.mark.parametrize("n,b", [(128, 64), (700, 37), (1000, 64), (5000, 100), (131, 1)])
def test_fast_scores_and_ranks_within_tolerance(model, n, b, cuda_device):
    ent, rel, heads, tails, rels = make_inputs(model, n, 128, b, seed=n + b)
    out, scores = _run_fast(model, ent, rel, heads, tails, rels, cuda_device)
    h, t, r = ent[heads].numpy(), ent[tails].numpy(), rel[rels].numpy()
    ref, mass = np_oracle.fast_mode_reference(model, ent.numpy(), h, t, r)
    got = scores.cpu().numpy().astype(np.float64)
    assert not np.isnan(got).any()
    err = np.abs(got - ref)
    assert (err <= TOL * mass + 1e-30).all(), f"max err/mass = {(err / np.maximum(mass, 1e-30)).max():.3e}"
    # the exact (bit-for-bit reference) path: true scores are shared, counts may differ only inside the band
    co = c_oracle.eval_rank(model, ent.numpy(), h, t, r, heads.numpy(), tails.numpy(), want_scores=True)
    assert np.array_equal(out["true_score"].reshape(-1).cpu().numpy(), co["true_score"])
    st = co["true_score"].astype(np.float64)[:, None]
    band = (np.abs(co["scores"].astype(np.float64) - st) <= 2 * TOL * mass).sum(1)       # includes the true entity
    for k in ("gt", "ge"):
        diff = np.abs(out[k].reshape(-1).cpu().numpy().astype(np.int64) - co[k].astype(np.int64))
        assert (diff <= band).all(), (k, int(diff.max()))
    assert bool((out["gt"] < out["ge"]).all())           # the true entity always ties itself


@pytest.mark.parametrize("model", FAST_MODELS)
def test_fast_matches_exact_mode_on_almost_all_queries(model, cuda_device):
    """FB15k-237-sized table: the two modes agree on the rank of nearly every query and on MRR."""
    n, b = 14541, 512
    ent, rel, heads, tails, rels = make_inputs(model, n, 128, b, seed=11, n_rel=237)
    dev = cuda_device
    triples = torch.stack([heads, tails, rels], dim=1).to(dev)
    exact = blp_b200.rank_sweep(model, ent.to(dev), rel.to(dev), triples)
    fast = blp_b200.rank_sweep(model, ent.to(dev), rel.to(dev), triples, mode="fast")
    assert torch.equal(exact["true_score"], fast["true_score"])
    same = ((exact["gt"] == fast["gt"]) & (exact["ge"] == fast["ge"])).float().mean().item()
    assert same >= 0.97, same
    assert int((exact["gt"] - fast["gt"]).abs().max()) <= 3
    me, mf = blp_b200.finalize(exact), blp_b200.finalize(fast)
    assert abs(me["mrr"] - mf["mrr"]) <= 1e-6
    for a, c in zip(me["hits_at_k"], mf["hits_at_k"]):
        assert abs(a - c) <= 2.0 / (2 * b)


@pytest.mark.parametrize("model", ("distmult", "complex"))
def test_fast_shard_sums_equal_full_table(model, cuda_device):
    n, b = 3000, 50
    ent, rel, heads, tails, rels = make_inputs(model, n, 128, b, seed=23)
    full, _ = _run_fast(model, ent, rel, heads, tails, rels, cuda_device, want_scores=False)
    acc = {k: torch.zeros_like(full[k]) for k in ("gt", "ge")}
    for rank in range(3):
        lo, hi = blp_b200.shard_bounds(n, 3, rank)
        part, _ = _run_fast(model, ent, rel, heads, tails, rels, cuda_device, want_scores=False, ent_offset=lo, shard=(lo, hi))
        for k in acc:
            acc[k] += part[k]
    for k in acc:
        assert torch.equal(acc[k], full[k]), k


def test_fast_mode_rejects_transe_and_other_widths(cuda_device):
    ent, rel, heads, tails, rels = make_inputs("transe", 300, 128, 4, seed=1)
    dev = cuda_device
    triples = torch.stack([heads, tails, rels], dim=1).to(dev)
    with pytest.raises(ValueError):
        blp_b200.rank_sweep("transe", ent.to(dev), rel.to(dev), triples, mode="fast")
    ent64 = torch.randn(300, 64)
    with pytest.raises(ValueError):
        ops.fast_table(ent64.to(dev))


@pytest.mark.parametrize("model", FAST_MODELS)
def test_fast_filtered_counters_stay_valid_with_tied_filtered_candidates(model, cuda_device):
    """the raw counters of the tensor-core mode come from split-FP16 scores, the filter correction re-scores
    with exact fp32.  Filtered candidates that TIE with the true score (known-true triples score near s_true; here exact
    duplicates of the true entity's row) may be judged differently by the two: the filtered counters must still be
    valid ranks (gt_f >= 0, ge_f >= gt_f + 1), never a zero / negative average rank."""
    n, b, n_rel = 400, 24, 5
    ent, rel, heads, tails, rels = make_inputs(model, n, 128, b, seed=77, n_rel=n_rel)
    heads, tails = heads % 300, tails % 300                # rows 300.. are the duplicates
    edges = []
    for i in range(b):
        dup_h, dup_t = 300 + i, 330 + i                    # duplicates of the true head / tail rows, both filtered
        ent[dup_h] = ent[heads[i]]
        ent[dup_t] = ent[tails[i]]
        edges += [(dup_h, int(tails[i]), int(rels[i])), (int(heads[i]), dup_t, int(rels[i]))]
    # make the true triples the best-scoring ones so that gt is tiny and a miscounted tie would go negative
    dev = cuda_device
    e, r = ent.to(dev), rel.to(dev)
    triples = torch.stack([heads, tails, rels], dim=1).to(dev)
    fidx = blp_b200.DeviceFilterIndex(np.array(edges, np.int64), None, n, n_rel, dev)
    exact = blp_b200.rank_sweep(model, e, r, triples, filter_index=fidx)
    fast = blp_b200.rank_sweep(model, e, r, triples, filter_index=fidx, mode="fast")
    for out in (exact, fast):
        gt_f, ge_f = out["gt_f"].cpu(), out["ge_f"].cpu()
        assert bool((gt_f >= 0).all()) and bool((ge_f >= gt_f + 1).all())
        assert bool(torch.isfinite(out["recip_f"]).all()) and float(out["recip_f"].max()) <= 1.0
    # exact mode: the duplicate is removed from the ties exactly
    assert torch.equal(exact["ge_f"], exact["ge"] - 1) and torch.equal(exact["gt_f"], exact["gt"])
    # fast mode: only candidates inside the tolerance band around s_true may be judged differently -- here the exact
    # duplicates of the true row (the filtered one, plus duplicates created for other triples with the same entity)
    ent_np = ent.numpy()
    for role, true_rows in ((0, heads), (1, tails)):
        band = np.array([(np.abs(ent_np - ent_np[int(t)]).max(axis=1) == 0).sum() - 1 for t in true_rows])
        assert bool(((fast["ge_f"][role].cpu() - exact["ge_f"][role].cpu()).abs().numpy() <= band).all())
        assert bool(((fast["gt_f"][role].cpu() - exact["gt_f"][role].cpu()).abs().numpy() <= band).all())


# ---- filter + refine: the tensor-core sweep with exact integer ranks (blp_rank_sweep_fast_exact) ----------------------
def _assert_same_counters(a, b, names=("gt", "ge")):
    assert torch.equal(a["true_score"], b["true_score"])
    for k in names:
        assert torch.equal(a[k], b[k]), (k, int((a[k] != b[k]).sum()))


@pytest.mark.parametrize("model", FAST_MODELS)
@pytest.mark.parametrize("n,b", [(128, 64), (700, 37), (1000, 64), (5000, 100), (131, 1), (14541, 300)])
def test_fast_exact_counters_equal_oracle(model, n, b, cuda_device):
    """gt / ge of the filter + refine mode are the oracle's integers (bit-exact ranks on the tensor path)."""
    ent, rel, heads, tails, rels = make_inputs(model, n, 128, b, seed=3 * n + b, n_rel=23)
    dev = cuda_device
    triples = torch.stack([heads, tails, rels], dim=1).to(dev)
    out = blp_b200.rank_sweep(model, ent.to(dev), rel.to(dev), triples, mode="fast_exact")
    assert "refine_overflow" not in out
    co = c_oracle.eval_rank(model, ent.numpy(), ent[heads].numpy(), ent[tails].numpy(), rel[rels].numpy(),
                            heads.numpy(), tails.numpy())
    assert np.array_equal(out["true_score"].reshape(-1).cpu().numpy(), co["true_score"])
    assert np.array_equal(out["gt"].reshape(-1).cpu().numpy(), co["gt"])
    assert np.array_equal(out["ge"].reshape(-1).cpu().numpy(), co["ge"])


@pytest.mark.parametrize("model,n,b", [("distmult", 14541, 2048), ("complex", 40943, 1024), ("simple", 14541, 1024)])
def test_fast_exact_equals_exact_mode_at_benchmark_sizes(model, n, b, cuda_device):
    """FB15k-237 / WN18RR-sized tables: every counter and metric of the two modes is identical, and the band the refine
    pass had to re-score is a handful of candidates per query."""
    ent, rel, heads, tails, rels = make_inputs(model, n, 128, b, seed=5, n_rel=237)
    dev = cuda_device
    e, r = ent.to(dev), rel.to(dev)
    triples = torch.stack([heads, tails, rels], dim=1).to(dev)
    exact = blp_b200.rank_sweep(model, e, r, triples)
    plan = blp_b200.RankSweepPlan(model, e, r, b, mode="fast_exact")
    fx = plan(triples)
    _assert_same_counters(exact, fx)
    assert torch.equal(exact["recip"].reshape(-1), fx["recip"].reshape(-1))
    assert torch.equal(exact["sums"][1:], fx["sums"][1:])                                  # hit counts
    assert abs(float(exact["sums"][0]) - float(fx["sums"][0])) <= 1e-12 * float(exact["sums"][0])   # fp64 sum order differs
    assert not plan.refine_overflowed()
    # worklist slots used (incl. the padded tails of the per-warp blocks: <= 148 x 8 x 128) per query; >= 1: the true
    # entity itself is always inside the band
    per_query = int(fx["refine_state"][0]) / (2 * b)
    assert 1.0 <= per_query <= 64.0 + 151552 / (2 * b), per_query
    # and through rank_sweep (chunked, its own worklist)
    fx2 = blp_b200.rank_sweep(model, e, r, triples, mode="fast_exact", chunk=300)
    _assert_same_counters(exact, fx2)


@pytest.mark.parametrize("model", FAST_MODELS)
def test_fast_exact_with_ties_duplicates_and_filters(model, cuda_device):
    """Duplicated rows tie with the true score bit for bit: they land in the band, the refine pass re-scores them with
    the reference's operations and the raw AND filtered counters equal the exact mode's (no clamping involved)."""
    n, b, n_rel = 400, 24, 5
    ent, rel, heads, tails, rels = make_inputs(model, n, 128, b, seed=78, n_rel=n_rel)
    heads, tails = heads % 300, tails % 300
    edges = []
    for i in range(b):
        ent[300 + i] = ent[heads[i]]
        ent[330 + i] = ent[tails[i]]
        edges += [(300 + i, int(tails[i]), int(rels[i])), (int(heads[i]), 330 + i, int(rels[i]))]
    ent[360:380] = 0.0                                       # zero rows: score exactly 0
    dev = cuda_device
    e, r = ent.to(dev), rel.to(dev)
    triples = torch.stack([heads, tails, rels], dim=1).to(dev)
    fidx = blp_b200.DeviceFilterIndex(np.array(edges, np.int64), None, n, n_rel, dev)
    exact = blp_b200.rank_sweep(model, e, r, triples, filter_index=fidx)
    fx = blp_b200.rank_sweep(model, e, r, triples, filter_index=fidx, mode="fast_exact")
    _assert_same_counters(exact, fx, ("gt", "ge", "gt_f", "ge_f"))
    assert torch.equal(exact["sums_f"][1:], fx["sums_f"][1:])
    assert abs(float(exact["sums_f"][0]) - float(fx["sums_f"][0])) <= 1e-12 * float(exact["sums_f"][0])


def test_fast_exact_shard_sums_equal_exact_full_table(cuda_device):
    model, n, b = "complex", 3000, 50
    ent, rel, heads, tails, rels = make_inputs(model, n, 128, b, seed=29)
    dev = cuda_device
    e, r = ent.to(dev), rel.to(dev)
    triples = torch.stack([heads, tails, rels], dim=1).to(dev)
    exact = blp_b200.rank_sweep(model, e, r, triples)
    acc = {k: torch.zeros_like(exact[k]) for k in ("gt", "ge")}
    for rank in range(3):
        lo, hi = blp_b200.shard_bounds(n, 3, rank)
        part = blp_b200.rank_sweep(model, e[lo:hi].contiguous(), r, triples, mode="fast_exact", ent_offset=lo,
                                   h_rows=e[triples[:, 0]], t_rows=e[triples[:, 1]])
        for k in acc:
            acc[k] += part[k]
    for k in acc:
        assert torch.equal(acc[k], exact[k]), k


def test_fast_exact_band_bounds_the_observed_error(cuda_device):
    """The a-priori band (kappa = 2e-5 of ||fold_abs|| * max||e||) against the measured |fast - reference| on random
    data: the observed maximum must stay below a quarter of it (margin for other data)."""
    for model in FAST_MODELS:
        n, b = 5000, 64
        ent, rel, heads, tails, rels = make_inputs(model, n, 128, b, seed=101)
        out, scores = _run_fast(model, ent, rel, heads, tails, rels, cuda_device)
        h, t, r = ent[heads].numpy(), ent[tails].numpy(), rel[rels].numpy()
        co = c_oracle.eval_rank(model, ent.numpy(), h, t, r, heads.numpy(), tails.numpy(), want_scores=True)
        err = np.abs(scores.cpu().numpy().astype(np.float64) - co["scores"].astype(np.float64)).max(axis=1)
        _, mass = np_oracle.fast_mode_reference(model, ent.numpy(), h, t, r)        # sum|terms| per (query, candidate)
        # the band is relative to a Cauchy-Schwarz bound >= mass.max(axis=1)
        assert (err <= 0.25 * 2e-5 * mass.max(axis=1) + 1e-30).all(), (model, float((err / mass.max(axis=1)).max()))


def test_fast_exact_overflow_falls_back_to_exact(cuda_device):
    """A table of identical rows puts EVERY candidate into the band: the worklist overflows, the flag is raised and
    rank_sweep redoes the sweep in exact mode (same results, `refine_overflow` reported)."""
    model, n, b = "distmult", 70000, 16
    ent, rel, heads, tails, rels = make_inputs(model, 64, 128, b, seed=7)
    ent = ent[:1].repeat(n, 1).contiguous()
    heads, tails = heads % n, tails % n
    dev = cuda_device
    e, r = ent.to(dev), rel.to(dev)
    triples = torch.stack([heads, tails, rels], dim=1).to(dev)
    exact = blp_b200.rank_sweep(model, e, r, triples)
    fx = blp_b200.rank_sweep(model, e, r, triples, mode="fast_exact")        # 2 * 16 * 70000 band entries > capacity 2^19
    assert fx.get("refine_overflow") is True
    _assert_same_counters(exact, fx)
    assert int(exact["ge"].min()) == n and int(exact["gt"].max()) == 0
