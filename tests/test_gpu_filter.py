"""Device-resident filter index (blp_filter_index_build / blp_filter_correct), MRR breakdowns
(blp_mrr_breakdown) and the device negative sampler (blp_negative_sample) -- SURVEY.md section 8f rows 1-3.

Filtered ranks are integers: bit-exact against the reference's eval_link_prediction (golden eval_loop_*)
and against the CSR path, which the other tests pin to utils.get_triple_filters' dense masks."""
import numpy as np
import pytest
import torch

from conftest import golden

import blp_b200
from blp_b200 import ops

pytestmark = pytest.mark.gpu
MODELS = ("transe", "distmult", "complex", "simple")


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


@pytest.mark.parametrize("model", MODELS)
def test_indexed_filter_reproduces_reference_eval_loop(model, cuda_device):
    """train.py:57-243 end to end with the filtering graph resident on the device."""
    g = golden("eval_loop_" + model)
    ent2idx = blp_b200.make_ent2idx(torch.from_numpy(g["entities"]), int(g["n_ids"]) - 1)
    triples = torch.from_numpy(g["triples"])
    rows = torch.stack([ent2idx[triples[:, 0]], ent2idx[triples[:, 1]], triples[:, 2]], dim=1)
    ent, rel = _t(g["ent_emb"], cuda_device), _t(g["rel_weight"], cuda_device)
    didx = blp_b200.DeviceFilterIndex(g["graph_edges"], ent2idx, ent.shape[0], rel.shape[0], cuda_device)
    out = blp_b200.rank_sweep(model, ent, rel, rows.to(cuda_device), filter_index=didx, chunk=40)
    # same integers as the host-CSR path (which mirrors utils.get_triple_filters mask by mask)
    fidx = blp_b200.TripleFilterIndex(g["graph_edges"], ent2idx)
    ref = blp_b200.rank_sweep(model, ent, rel, rows.to(cuda_device), filter_index=fidx, filter_triples=g["triples"], chunk=40)
    for k in ("gt", "ge", "gt_f", "ge_f"):
        assert torch.equal(out[k], ref[k]), k
    m = blp_b200.finalize(out)
    want = dict(zip(g["scalar_names"].tolist(), g["scalar_values"].tolist()))
    assert abs(m["mrr_f"] - want["test_mrr_filt"]) <= 1e-6
    for j, k in enumerate((1, 3, 10)):
        assert abs(m["hits_at_k_f"][j] - want[f"test_hits@{k}_filt"]) <= 1e-9
    # train.py:173-188, 215-225: by-position breakdown of the filtered MRR
    bd = blp_b200.breakdowns(out, g["triples"], new_entities=set(g["new_entities"].tolist()),
                             rel_categories=g["rel_categories"], max_ent_id=int(g["n_ids"]) - 1)
    cnt = bd["mrr_pos_counts"].cpu().numpy().copy()
    cnt[cnt < 1.0] = 1.0
    by_pos = bd["mrr_by_position"].cpu().numpy() / cnt
    for i, name in enumerate(("test_mrr_filt_both_new", "test_mrr_filt_head_new", "test_mrr_filt_tail_new")):
        assert abs(by_pos[i] - want[name]) <= 2e-6, (name, by_pos[i], want[name])
    # utils.split_by_category (utils.py:150-168) restated with numpy on the same reciprocal ranks
    recip = out["recip_f"].cpu().numpy().reshape(2, -1).astype(np.float64)
    cats = g["rel_categories"][g["triples"][:, 2]]
    want_cat = np.zeros((2, 4)); want_cnt = np.zeros(4)
    for c in range(4):
        want_cat[0, c] = recip[0, cats == c].sum(); want_cat[1, c] = recip[1, cats == c].sum()
        want_cnt[c] = (cats == c).sum()
    assert np.allclose(bd["mrr_by_category"].cpu().numpy(), want_cat, rtol=0, atol=1e-9)
    assert np.array_equal(bd["mrr_cat_count"].cpu().numpy().reshape(-1), want_cnt)


@pytest.mark.parametrize("model", ("transe", "complex"))
@pytest.mark.parametrize("shards", (1, 3))
def test_indexed_filter_dense_graph_with_parallel_edges_and_missing_rows(model, shards, cuda_device):
    """N-to-N relations (tens of filtered candidates per query), duplicate edges, entities without a table row
    (ent2idx == -1), test triples whose (head, rel) never occurs; candidate axis split over `shards` row blocks."""
    rng = np.random.default_rng(7)
    n_ids, n_rows, n_rel, d, T = 400, 300, 3, 128, 130
    g = torch.Generator().manual_seed(11)
    ent = torch.randn(n_rows, d, generator=g)
    if model == "transe":
        ent = torch.nn.functional.normalize(ent, dim=-1)
    rel = (torch.rand(n_rel, d, generator=g) * 2 - 1) * 0.2
    entities = torch.from_numpy(rng.permutation(n_ids)[:n_rows].astype(np.int64))
    ent2idx = blp_b200.make_ent2idx(entities, n_ids - 1)
    edges = np.stack([rng.integers(0, n_ids, 6000), rng.integers(0, n_ids, 6000), rng.integers(0, n_rel, 6000)], 1)
    edges = np.concatenate([edges, edges[:500]])                                   # parallel edges
    tr_ids = np.stack([entities.numpy()[rng.integers(0, n_rows, T)], entities.numpy()[rng.integers(0, n_rows, T)],
                       rng.integers(0, n_rel, T)], 1)
    edges = np.concatenate([edges, tr_ids[: T // 2]])                              # half of the test triples are graph edges
    rows = torch.stack([ent2idx[torch.from_numpy(tr_ids[:, 0])], ent2idx[torch.from_numpy(tr_ids[:, 1])],
                        torch.from_numpy(tr_ids[:, 2])], dim=1).to(cuda_device)
    e, r = ent.to(cuda_device), rel.to(cuda_device)
    fidx = blp_b200.TripleFilterIndex(edges, ent2idx)
    ref = blp_b200.rank_sweep(model, e, r, rows, filter_index=fidx, filter_triples=tr_ids, chunk=64)
    assert int((ref["gt"] - ref["gt_f"]).sum()) > 0                                 # the filter does remove candidates
    didx = blp_b200.DeviceFilterIndex(edges, ent2idx, n_rows, n_rel, cuda_device)
    tot = None
    for s in range(shards):
        lo, hi = blp_b200.shard_bounds(n_rows, shards, s)
        part = blp_b200.rank_sweep(model, e[lo:hi].contiguous(), r, rows, filter_index=didx, ent_offset=lo, chunk=50,
                                   h_rows=e[rows[:, 0]], t_rows=e[rows[:, 1]])
        cur = {k: part[k].clone() for k in ("gt", "ge", "gt_f", "ge_f")}
        tot = cur if tot is None else {k: tot[k] + cur[k] for k in cur}
    for k in ("gt", "ge", "gt_f", "ge_f"):
        assert torch.equal(tot[k], ref[k]), k


def test_indexed_filter_identity_ids_and_empty_graph(cuda_device):
    """Transductive models: ids are rows (ent2idx=None, train.py:89-93); an empty graph filters nothing."""
    model, n, d, T = "distmult", 200, 128, 33
    g = torch.Generator().manual_seed(3)
    ent, rel = torch.randn(n, d, generator=g).to(cuda_device), (torch.rand(4, d, generator=g) * 0.2).to(cuda_device)
    rows = torch.stack([torch.randint(0, n, (T,), generator=g), torch.randint(0, n, (T,), generator=g),
                        torch.randint(0, 4, (T,), generator=g)], dim=1)
    edges = np.stack([np.arange(n).repeat(5), np.tile(np.arange(n), 5), np.zeros(5 * n, np.int64)], 1)
    didx = blp_b200.DeviceFilterIndex(edges, None, n, 4, cuda_device)
    out = blp_b200.rank_sweep(model, ent, rel, rows.to(cuda_device), filter_index=didx)
    fidx = blp_b200.TripleFilterIndex(edges, np.arange(n))
    ref = blp_b200.rank_sweep(model, ent, rel, rows.to(cuda_device), filter_index=fidx, filter_triples=rows.numpy())
    for k in ("gt_f", "ge_f"):
        assert torch.equal(out[k], ref[k]), k
    empty = blp_b200.DeviceFilterIndex(np.zeros((0, 3), np.int64), None, n, 4, cuda_device)
    out0 = blp_b200.rank_sweep(model, ent, rel, rows.to(cuda_device), filter_index=empty)
    assert torch.equal(out0["gt_f"], out0["gt"]) and torch.equal(out0["ge_f"], out0["ge"])


# ------------------------------------------------------------------ negative sampler ----
def test_negative_sampler_layout_matches_reference(cuda_device):
    """Shape / dtype / strides of data.get_negative_sampling_indices (kat.npz: (4,3,2), stride (2,8,1))."""
    kat = golden("kat")
    neg = blp_b200.get_negative_sampling_indices(4, 3, device=cuda_device, seed=1)
    assert tuple(neg.shape) == tuple(kat["neg_idx_shape"]) and tuple(neg.stride()) == tuple(kat["neg_idx_stride"])
    assert neg.dtype == torch.int64 and not neg.is_contiguous()
    neg = blp_b200.get_negative_sampling_indices(32, 64, repeats=2, device=cuda_device, seed=1)
    assert tuple(neg.shape) == (64, 64, 2) and tuple(neg.stride()) == (2, 128, 1)
    assert int(neg.min()) >= 0 and int(neg.max()) <= 63


@pytest.mark.parametrize("b,k,repeats", [(64, 512, 1), (8, 2000, 2), (2, 4096, 1)])
def test_negative_sampler_structure_and_distribution(b, k, repeats, cuda_device):
    """data.py:35-81: exactly one column keeps the row's own entity, the other is uniform over the 2B - 2
    entities of the other rows; the corrupted side is a fair coin."""
    neg = blp_b200.get_negative_sampling_indices(b, k, repeats, device=cuda_device, seed=1234, offset=5).cpu().numpy()
    rows = np.arange(b * repeats) % b
    own0, own1 = (2 * rows)[:, None], (2 * rows + 1)[:, None]
    keep_head, keep_tail = neg[:, :, 0] == own0, neg[:, :, 1] == own1
    assert np.all(keep_head ^ keep_tail)                       # exactly one side is corrupted
    repl = np.where(keep_head, neg[:, :, 1], neg[:, :, 0])
    assert repl.min() >= 0 and repl.max() < 2 * b
    assert not np.any((repl == own0) | (repl == own1))         # never an entity of the own row
    n = repl.size
    frac_tail = keep_head.mean()
    assert abs(frac_tail - 0.5) < 5.0 * 0.5 / np.sqrt(n)
    if b > 2:
        # chi-square of the replacement index against uniform over the 2B - 2 allowed values, pooled over rows
        shifted = repl - 2 * (repl > own1)                     # map the allowed values of each row onto [0, 2B-2)
        counts = np.bincount(shifted.reshape(-1), minlength=2 * b - 2).astype(np.float64)
        exp = n / (2 * b - 2)
        chi2 = ((counts - exp) ** 2 / exp).sum()
        dof = 2 * b - 3
        assert chi2 < dof + 6.0 * np.sqrt(2.0 * dof), (chi2, dof)
    # a different offset / seed gives a different stream; the same (seed, offset) reproduces
    again = blp_b200.get_negative_sampling_indices(b, k, repeats, device=cuda_device, seed=1234, offset=5).cpu().numpy()
    other = blp_b200.get_negative_sampling_indices(b, k, repeats, device=cuda_device, seed=1234, offset=6).cpu().numpy()
    assert np.array_equal(neg, again) and not np.array_equal(neg, other)


def test_negative_sampler_feeds_compute_loss_in_place(cuda_device):
    """The sampler's strided output is consumed by the fused loss without a copy, like the reference's (models.py:65)."""
    from oracle import c_oracle
    b, k, d = 16, 48, 128
    g = torch.Generator().manual_seed(0)
    m = blp_b200.TransductiveLinkPrediction(d, "transe", "margin", 50, 5, 0).to(cuda_device)
    pairs, rels = torch.randint(0, 50, (b, 2), generator=g), torch.randint(0, 5, (b, 1), generator=g)
    neg = blp_b200.get_negative_sampling_indices(b, k, device=cuda_device, seed=9)
    ent_embs = m.encode(pairs.to(cuda_device)).detach().requires_grad_(True)
    loss = m.compute_loss(ent_embs, rels.to(cuda_device), neg)
    loss.backward()
    co = c_oracle.train_loss("transe", "margin", ent_embs.detach().cpu().numpy(),
                             m.rel_emb.weight.detach().cpu()[rels[:, 0]].numpy(), neg.cpu().numpy(), 0.0)
    assert abs(loss.item() - float(co["loss"])) <= 1e-5 * abs(float(co["loss"]))
    assert np.abs(ent_embs.grad.cpu().numpy() - co["grad_ent"]).max() <= 2e-5 * np.abs(co["grad_ent"]).max()
    with pytest.raises(ValueError):
        blp_b200.get_negative_sampling_indices(1, 4, device=cuda_device)            # data.py:289-291 needs batch > 1


@pytest.mark.parametrize("model,mode", [("transe", "exact"), ("complex", "exact"), ("distmult", "fast")])
def test_rank_sweep_plan_equals_rank_sweep(model, mode, cuda_device):
    """RankSweepPlan = rank_sweep with the per-call host work hoisted: same integers, same metrics, for fresh
    triples on every call, raw and filtered."""
    g = golden("eval_loop_" + ("distmult" if model == "distmult" else model))
    ent2idx = blp_b200.make_ent2idx(torch.from_numpy(g["entities"]), int(g["n_ids"]) - 1)
    triples = torch.from_numpy(g["triples"])
    rows = torch.stack([ent2idx[triples[:, 0]], ent2idx[triples[:, 1]], triples[:, 2]], dim=1).to(cuda_device)
    ent, rel = _t(g["ent_emb"], cuda_device), _t(g["rel_weight"], cuda_device)
    didx = blp_b200.DeviceFilterIndex(g["graph_edges"], ent2idx, ent.shape[0], rel.shape[0], cuda_device)
    T = 32
    plan = blp_b200.RankSweepPlan(model, ent, rel, T, mode=mode, filter_index=didx)
    for lo in (0, 32, 64):
        chunk = rows[lo:lo + T].contiguous()
        got = plan(chunk)
        want = blp_b200.rank_sweep(model, ent, rel, chunk, filter_index=didx, mode=mode)
        for k in ("gt", "ge", "gt_f", "ge_f", "true_score", "recip", "recip_f", "hits", "hits_f", "sums", "sums_f"):
            assert torch.equal(got[k], want[k]), (k, lo)
    with pytest.raises(ValueError):
        plan(rows[:T + 1].contiguous())


# ------------------------------------------------------- entity-table production (row f4) ----
@pytest.mark.parametrize("d", (128, 300, 768, 100))
def test_store_rows_normalize_matches_reference_encode(d, cuda_device):
    """blp_store_rows(normalize=1) == the reference's encode() bits (golden normalize.npz), incl. zero / tiny rows."""
    g = golden("normalize")
    x, y = g[f"x_{d}"], g[f"y_{d}"]
    shard = torch.full((x.shape[0], d), float("nan"), device=cuda_device)
    blp_b200.store_rows(shard, _t(x, cuda_device), normalize=True)
    assert np.array_equal(shard.cpu().numpy(), y)


@pytest.mark.parametrize("d", (67, 1030, 2048))
def test_store_rows_normalize_other_widths(d, cuda_device):
    """Widths outside the vectorised path (odd d, d > 1024) against the oracle restatement of F.normalize."""
    from oracle import np_oracle
    x = torch.randn(33, d, generator=torch.Generator().manual_seed(d))
    shard = torch.empty((33, d), device=cuda_device)
    blp_b200.store_rows(shard, x.to(cuda_device), normalize=True)
    assert np.array_equal(shard.cpu().numpy(), np_oracle.l2_normalize_rows(x.numpy()))


def test_store_rows_scatters_into_row_shards(cuda_device):
    """train.py:95-123 with the table row-sharded: every rank is handed every encoder batch and keeps its rows;
    the union of the shards is the reference's dense table, and the sweep over the shards equals the dense sweep."""
    from oracle import np_oracle
    n, d, world, bs = 1000, 128, 3, 96
    g = torch.Generator().manual_seed(4)
    raw = torch.randn(n, d, generator=g)
    dense = torch.from_numpy(np_oracle.l2_normalize_rows(raw.numpy()))                 # == F.normalize(raw)
    shards = []
    for rank in range(world):
        lo, hi = blp_b200.shard_bounds(n, world, rank)
        shard = torch.zeros((hi - lo, d), device=cuda_device)
        for idx in range(0, n, bs):                                                     # the reference's encode loop
            blp_b200.store_rows(shard, raw[idx:idx + bs].to(cuda_device), row0=idx, normalize=True, ent_offset=lo)
        shards.append(shard)
    assert torch.equal(torch.cat(shards).cpu(), dense)
    # permuted destinations (entities[idx] order) and no normalisation (bilinear models, models.py:18)
    perm = torch.randperm(n, generator=g)
    full = torch.zeros((n, d), device=cuda_device)
    blp_b200.store_rows(full, raw.to(cuda_device), rows=perm.to(cuda_device))
    assert torch.equal(full.cpu()[perm], raw)


@pytest.mark.parametrize("normalize", (True, False))
def test_store_rows_stream_path(normalize, cuda_device):
    """Bulk table production (>= 4,096 contiguous rows of d = 128): the persistent TMA pipeline (store_rows_stream_kernel:
    bulk loads -> in-place normalisation in shared memory -> bulk stores).  Bit-equal to F.normalize's CPU bits (the
    oracle restatement pinned on the reference's encode()), ragged last tile, zero / tiny rows, and row-sharded
    destinations clipped on the host (rows owned by other ranks are neither read nor written)."""
    from oracle import np_oracle
    n, d = 40003, 128
    g = torch.Generator().manual_seed(12)
    raw = torch.randn(n, d, generator=g)
    raw[5] = 0.0
    raw[77] *= 1e-30
    raw[40002] *= 1e20
    want = torch.from_numpy(np_oracle.l2_normalize_rows(raw.numpy())) if normalize else raw
    x = raw.to(cuda_device)
    full = torch.full((n, d), float("nan"), device=cuda_device)
    blp_b200.store_rows(full, x, normalize=normalize)
    assert torch.equal(full.cpu(), want)
    # three row shards, encoder batches of 8,192 rows handed to every rank (train.py:95-123)
    for rank in range(3):
        lo, hi = blp_b200.shard_bounds(n, 3, rank)
        shard = torch.full((hi - lo, d), float("nan"), device=cuda_device)
        for idx in range(0, n, 8192):
            blp_b200.store_rows(shard, x[idx:idx + 8192], row0=idx, normalize=normalize, ent_offset=lo)
        assert torch.equal(shard.cpu(), want[lo:hi]), rank


@pytest.mark.parametrize("b,seed", [(64, 0), (7, 1), (1, 2)])
def test_split_by_functions_equal_the_reference(b, seed, cuda_device):
    """utils.split_by_new_position / utils.split_by_category (utils.py:114-168) as patch() rebinds them: one launch per
    call instead of per-triple Python loops; same (3,), (3,), (2, 4), (1, 4) fp32 tensors as the reference's own
    functions (run here on the CPU from oracle/_ref or /root/reference) -- counts exactly, sums to fp32 rounding."""
    from oracle import ref_loader
    if ref_loader.available() is None:
        pytest.skip("reference modules not built (python oracle/build_ref.py)")
    ref_utils = ref_loader.load(("models", "utils"))["utils"]       # utils.py imports models
    g = torch.Generator().manual_seed(seed)
    n_ids, n_rel = 500, 11
    triples = torch.stack([torch.randint(0, n_ids, (b,), generator=g), torch.randint(0, n_ids, (b,), generator=g),
                           torch.randint(0, n_rel, (b,), generator=g)], dim=1)
    recip = 1.0 / torch.randint(1, 200, (2 * b,), generator=g).float()
    new_entities = set(torch.randint(0, n_ids, (200,), generator=g).tolist())
    rel_categories = torch.randint(0, 4, (n_rel,), generator=g)
    want_pos, want_cnt = ref_utils.split_by_new_position(triples, recip, new_entities)
    want_cat, want_ccnt = ref_utils.split_by_category(triples, recip, rel_categories)
    dev = cuda_device
    for _ in range(2):                                       # second round: the cached membership mask
        got_pos, got_cnt = blp_b200.split_by_new_position(triples, recip.to(dev), new_entities)
        got_cat, got_ccnt = blp_b200.split_by_category(triples, recip.to(dev), rel_categories.to(dev))
        assert got_pos.dtype == torch.float32 and got_pos.device.type == "cuda" and tuple(got_pos.shape) == (3,)
        assert tuple(got_cat.shape) == (2, 4) and tuple(got_ccnt.shape) == (1, 4)
        assert torch.equal(got_cnt.cpu(), want_cnt) and torch.equal(got_ccnt.cpu(), want_ccnt)
        assert torch.allclose(got_pos.cpu(), want_pos, rtol=2e-6, atol=0)
        assert torch.allclose(got_cat.cpu(), want_cat, rtol=2e-6, atol=0)
    new_entities.add(n_ids + 5)                               # a changed set is re-read
    t2 = torch.tensor([[n_ids + 5, 0, 0]])
    pos, cnt = blp_b200.split_by_new_position(t2, torch.tensor([0.5, 0.25], device=dev), new_entities)
    assert cnt.sum().item() == 1.0
