"""CUDA eval path vs the oracle and the golden vectors (through the C ABI).
Bit-exact: every score, every integer rank counter, reciprocal ranks and hits."""
import numpy as np
import pytest
import torch

from conftest import golden, golden_names, mask_to_csr, unpack_mask
from oracle import c_oracle

import blp_b200
from blp_b200 import ops

pytestmark = pytest.mark.gpu
EVAL = [n for n in golden_names("eval_") if not n.startswith("eval_loop")]
MODELS = ("transe", "distmult", "complex", "simple")


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def make_inputs(model, n, d, b, seed, n_rel=11):
    g = torch.Generator().manual_seed(seed)
    ent = torch.randn(n, d, generator=g)
    if model == "transe":
        ent = torch.nn.functional.normalize(ent, dim=-1)
    a = (6.0 / (n_rel + d)) ** 0.5
    rel = (torch.rand(n_rel, d, generator=g) * 2 - 1) * a
    heads = torch.randint(0, n, (b,), generator=g)
    tails = torch.randint(0, n, (b,), generator=g)
    rels = torch.randint(0, n_rel, (b,), generator=g)
    return ent, rel, heads, tails, rels


@pytest.mark.parametrize("name", EVAL)
def test_golden_eval_rank(name, cuda_device):
    g = golden(name)
    model = name.split("_")[1]
    heads, tails, rels = g["heads"][:, 0], g["tails"][:, 0], g["rels"][:, 0]
    indptr, idx = mask_to_csr(unpack_mask(g))
    ent = _t(g["ent"], cuda_device)
    out = ops.eval_rank(model, ent, _t(g["ent"][heads], cuda_device), _t(g["ent"][tails], cuda_device),
                        _t(g["rel"][rels], cuda_device), _t(indptr, cuda_device),
                        _t(idx if len(idx) else np.zeros(1, np.int64), cuda_device))
    for k in ("gt", "ge", "gt_f", "ge_f"):
        assert np.array_equal(out[k].cpu().numpy(), g[k]), k
    true = np.concatenate([heads, tails])
    assert np.array_equal(out["true_score"].cpu().numpy(), g["pred"][np.arange(len(true)), true])
    recip, hits = ops.metrics_from_counts(out["gt"], out["ge"], [1, 3, 10])
    assert np.array_equal(recip.cpu().numpy(), g["recip"]) and np.array_equal(hits.cpu().numpy(), g["hits"])
    recip, hits = ops.metrics_from_counts(out["gt_f"], out["ge_f"], [1, 3, 10])
    assert np.array_equal(recip.cpu().numpy(), g["recip_f"]) and np.array_equal(hits.cpu().numpy(), g["hits_f"])
    sums = ops.metrics_reduce(out["gt"], out["ge"], [1, 3, 10]).cpu().numpy()
    assert abs(sums[0] - g["recip"].astype(np.float64).sum()) < 1e-9
    assert np.array_equal(sums[1:], g["hits"].sum(0).astype(np.float64))


@pytest.mark.parametrize("name", EVAL)
def test_golden_score_fn_and_get_metrics(name, cuda_device):
    """The reference's own call sequence (train.py:141-153) through the drop-in functions."""
    g = golden(name)
    model = name.split("_")[1]
    fn = getattr(blp_b200, model + "_score")
    ent = _t(g["ent"], cuda_device)
    heads, tails, rels = (_t(g[k], cuda_device) for k in ("heads", "tails", "rels"))
    ent_emb = ent.unsqueeze(0)
    head_embs, tail_embs = ent[heads], ent[tails]              # (B,1,D)
    rel_embs = _t(g["rel"], cuda_device)[rels]
    hp = fn(ent_emb, tail_embs, rel_embs)
    tp = fn(head_embs, ent_emb, rel_embs)
    pred = torch.cat((hp, tp))
    assert np.array_equal(pred.cpu().numpy(), g["pred"])
    true = torch.cat((heads, tails))
    recip, hits = blp_b200.get_metrics(pred, true, torch.tensor([[1, 3, 10]], device=cuda_device))
    assert recip.shape == (pred.shape[0], 1) and hits.dtype == torch.bool
    assert np.array_equal(recip.cpu().numpy(), g["recip"]) and np.array_equal(hits.cpu().numpy(), g["hits"])
    # filtered re-rank exactly as train.py:164-167 writes it
    mask = _t(unpack_mask(g), cuda_device)
    pred[mask] = pred.min() - 1.0
    recip, hits = blp_b200.get_metrics(pred, true, torch.tensor([[1, 3, 10]], device=cuda_device))
    assert np.array_equal(recip.cpu().numpy(), g["recip_f"]) and np.array_equal(hits.cpu().numpy(), g["hits_f"])


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("n,b", [(1, 1), (31, 2), (257, 3), (128, 64), (129, 65), (1000, 7), (777, 12), (300, 33),
                                 (14541, 64)])
def test_sweep_vs_oracle_d128(model, n, b, cuda_device):
    ent, rel, heads, tails, rels = make_inputs(model, n, 128, b, seed=n * 7 + b)
    if n > 40:
        ent[n - 1] = ent[3]            # exact ties
        heads[0] = 3
    co = c_oracle.eval_rank(model, ent.numpy(), ent[heads].numpy(), ent[tails].numpy(), rel[rels].numpy(),
                            heads.numpy(), tails.numpy())
    e = ent.to(cuda_device)
    out = ops.eval_rank(model, e, e[heads.to(cuda_device)], e[tails.to(cuda_device)], rel[rels].to(cuda_device))
    assert np.array_equal(out["true_score"].cpu().numpy(), co["true_score"])
    assert np.array_equal(out["gt"].cpu().numpy(), co["gt"])
    assert np.array_equal(out["ge"].cpu().numpy(), co["ge"])


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("d", (2, 64, 100, 256, 300, 768))
def test_generic_width_vs_oracle(model, d, cuda_device):
    n, b = 203, 5
    ent, rel, heads, tails, rels = make_inputs(model, n, d, b, seed=d)
    co = c_oracle.eval_rank(model, ent.numpy(), ent[heads].numpy(), ent[tails].numpy(), rel[rels].numpy(),
                            heads.numpy(), tails.numpy(), want_scores=True)
    e = ent.to(cuda_device)
    out = ops.eval_rank(model, e, e[heads.to(cuda_device)], e[tails.to(cuda_device)], rel[rels].to(cuda_device))
    assert np.array_equal(out["true_score"].cpu().numpy(), co["true_score"])
    assert np.array_equal(out["gt"].cpu().numpy(), co["gt"]) and np.array_equal(out["ge"].cpu().numpy(), co["ge"])
    fn = getattr(blp_b200, model + "_score")
    r = rel[rels].to(cuda_device).unsqueeze(1)
    hp = fn(e.unsqueeze(0), e[tails.to(cuda_device)].unsqueeze(1), r)
    tp = fn(e[heads.to(cuda_device)].unsqueeze(1), e.unsqueeze(0), r)
    assert np.array_equal(torch.cat((hp, tp)).cpu().numpy(), co["scores"])


def test_odd_width_rejected_for_halves_models(cuda_device):
    x = torch.zeros(2, 1, 7, device=cuda_device)
    for name in ("complex_score", "simple_score"):
        with pytest.raises(ValueError):
            getattr(blp_b200, name)(x, x, x)


def test_empty_inputs(cuda_device):
    ent = torch.randn(10, 128, device=cuda_device)
    empty = torch.empty(0, 128, device=cuda_device)
    out = ops.eval_rank("transe", ent, empty, empty, empty)
    assert out["gt"].numel() == 0
    q = torch.randn(3, 128, device=cuda_device)
    out = ops.eval_rank("distmult", torch.empty(0, 128, device=cuda_device), q, q, q)
    assert out["gt"].tolist() == [0] * 6 and out["ge"].tolist() == [0] * 6


@pytest.mark.parametrize("model", MODELS)
def test_shard_sums_equal_full_table(model, cuda_device):
    """Integer counters add across row shards: any partition gives identical ranks."""
    n, b = 5000, 40
    ent, rel, heads, tails, rels = make_inputs(model, n, 128, b, seed=99)
    e = ent.to(cuda_device)
    h, t, r = e[heads.to(cuda_device)], e[tails.to(cuda_device)], rel[rels].to(cuda_device)
    mask = torch.rand(2 * b, n) < 0.02
    mask[torch.arange(2 * b), torch.cat([heads, tails])] = False
    indptr, idx = mask_to_csr(mask.numpy())
    ip, ix = _t(indptr, cuda_device), _t(idx, cuda_device)
    full = ops.eval_rank(model, e, h, t, r, ip, ix)
    for world in (2, 3, 8):
        acc = {k: torch.zeros_like(full[k]) for k in ("gt", "ge", "gt_f", "ge_f")}
        for rank in range(world):
            lo, hi = blp_b200.shard_bounds(n, world, rank)
            part = ops.eval_rank(model, e[lo:hi], h, t, r, ip, ix, ent_offset=lo)
            assert torch.equal(part["true_score"], full["true_score"])
            for k in acc:
                acc[k] += part[k]
        for k in acc:
            assert torch.equal(acc[k], full[k]), (k, world)
    co = c_oracle.eval_rank(model, ent.numpy(), ent[heads].numpy(), ent[tails].numpy(), rel[rels].numpy(),
                            heads.numpy(), tails.numpy(), indptr, idx)
    for k in ("gt", "ge", "gt_f", "ge_f"):
        assert np.array_equal(full[k].cpu().numpy(), co[k]), k


@pytest.mark.parametrize("model", MODELS)
def test_rank_sweep_reproduces_reference_eval_loop(model, cuda_device):
    """train.py:57-243 end to end (golden eval_loop_*): raw + filtered MRR and hits@k."""
    g = golden("eval_loop_" + model)
    ent2idx = blp_b200.make_ent2idx(torch.from_numpy(g["entities"]), int(g["n_ids"]) - 1)
    triples = torch.from_numpy(g["triples"])
    rows = torch.stack([ent2idx[triples[:, 0]], ent2idx[triples[:, 1]], triples[:, 2]], dim=1)
    fidx = blp_b200.TripleFilterIndex(g["graph_edges"], ent2idx)
    out = blp_b200.rank_sweep(model, _t(g["ent_emb"], cuda_device), _t(g["rel_weight"], cuda_device),
                              rows.to(cuda_device), filter_index=fidx, filter_triples=g["triples"], chunk=40)
    m = blp_b200.finalize(out)
    want = dict(zip(g["scalar_names"].tolist(), g["scalar_values"].tolist()))
    assert abs(m["mrr"] - want["test_mrr"]) <= 1e-6 and abs(m["mrr_f"] - want["test_mrr_filt"]) <= 1e-6
    for j, k in enumerate((1, 3, 10)):
        assert abs(m["hits_at_k"][j] - want[f"test_hits@{k}"]) <= 1e-9
        assert abs(m["hits_at_k_f"][j] - want[f"test_hits@{k}_filt"]) <= 1e-9


def test_full_size_properties_fb15k237(cuda_device):
    """BASELINE config 2 at full size (N=14,541, 2,048 triples): size-independent checks.
    (1) shard sums == full table; (2) permuting candidate rows leaves every rank unchanged;
    (3) gt < ge (the true entity always ties itself); (4) a sampled subset of queries equals the oracle."""
    model, n, b = "transe", 14541, 2048
    ent, rel, heads, tails, rels = make_inputs(model, n, 128, b, seed=5, n_rel=237)
    e = ent.to(cuda_device)
    hd, td = heads.to(cuda_device), tails.to(cuda_device)
    r = rel[rels].to(cuda_device)
    full = ops.eval_rank(model, e, e[hd], e[td], r)
    assert bool((full["gt"] < full["ge"]).all())
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(1)).to(cuda_device)
    shuffled = ops.eval_rank(model, e[perm], e[hd], e[td], r)
    assert torch.equal(shuffled["gt"], full["gt"]) and torch.equal(shuffled["ge"], full["ge"])
    lo, hi = blp_b200.shard_bounds(n, 2, 1)
    a = ops.eval_rank(model, e[:lo], e[hd], e[td], r)
    c = ops.eval_rank(model, e[lo:hi], e[hd], e[td], r, ent_offset=lo)
    assert torch.equal(a["gt"] + c["gt"], full["gt"]) and torch.equal(a["ge"] + c["ge"], full["ge"])
    sel = torch.arange(0, b, 64)
    co = c_oracle.eval_rank(model, ent.numpy(), ent[heads[sel]].numpy(), ent[tails[sel]].numpy(),
                            rel[rels[sel]].numpy(), heads[sel].numpy(), tails[sel].numpy())
    qsel = torch.cat([sel, sel + b])
    assert np.array_equal(full["gt"].cpu()[qsel].numpy(), co["gt"])
    assert np.array_equal(full["ge"].cpu()[qsel].numpy(), co["ge"])


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("d,chunk", [(128, 1000), (128, 17), (64, 23)])
def test_native_sweep_gathers_in_kernel(model, d, chunk, cuda_device):
    """blp_rank_sweep (train.py:141-143 gathers folded into the kernels, (2, T) outputs, chunk-local CSR)
    equals blp_eval_rank on pre-gathered rows and the oracle, for any chunking."""
    n, T = 700, 50
    ent, rel, heads, tails, rels = make_inputs(model, n, d, T, seed=31 + d)
    mask = torch.rand(2 * T, n) < 0.03
    mask[torch.arange(2 * T), torch.cat([heads, tails])] = False
    indptr, idx = mask_to_csr(mask.numpy())
    co = c_oracle.eval_rank(model, ent.numpy(), ent[heads].numpy(), ent[tails].numpy(), rel[rels].numpy(),
                            heads.numpy(), tails.numpy(), indptr, idx)
    triples = torch.stack([heads, tails, rels], dim=1).to(cuda_device)
    out = blp_b200.rank_sweep(model, ent.to(cuda_device), rel.to(cuda_device), triples, filter_csr=(indptr, idx), chunk=chunk)
    for k in ("gt", "ge", "gt_f", "ge_f"):
        assert out[k].shape == (2, T)
        assert np.array_equal(out[k].reshape(-1).cpu().numpy(), co[k]), k
    assert np.array_equal(out["true_score"].reshape(-1).cpu().numpy(), co["true_score"])
    # fused metrics kernel == the two-step form
    recip, hits = ops.metrics_from_counts(out["gt_f"].reshape(-1), out["ge_f"].reshape(-1), [1, 3, 10])
    assert torch.equal(recip, out["recip_f"]) and torch.equal(hits, out["hits_f"])
    sums = ops.metrics_reduce(out["gt"].reshape(-1), out["ge"].reshape(-1), [1, 3, 10])
    assert torch.equal(sums, out["sums"])


def test_native_sweep_out_of_range_ids_are_flagged(cuda_device):
    """train.py:137-138 asserts mapped ids are >= 0; here such a triple gets a NaN true score and zero counts."""
    n, T = 300, 6
    ent, rel, heads, tails, rels = make_inputs("transe", n, 128, T, seed=3)
    triples = torch.stack([heads, tails, rels], dim=1)
    triples[2, 0] = -1
    triples[4, 2] = 99
    out = blp_b200.rank_sweep("transe", ent.to(cuda_device), rel.to(cuda_device), triples.to(cuda_device))
    ts = out["true_score"].cpu()
    bad = torch.isnan(ts[0])
    assert bad.tolist() == [False, False, True, False, True, False]
    assert out["ge"].cpu()[:, bad].sum() == 0 and bool((out["ge"].cpu()[:, ~bad] >= 1).all())


@pytest.mark.parametrize("model", ("transe", "complex"))
def test_native_sweep_sharded_with_pregathered_rows(model, cuda_device):
    """Entity-sharded form: every shard gets the replicated true rows; integer counters add up exactly."""
    n, T = 3000, 70
    ent, rel, heads, tails, rels = make_inputs(model, n, 128, T, seed=17)
    e, rw = ent.to(cuda_device), rel.to(cuda_device)
    triples = torch.stack([heads, tails, rels], dim=1).to(cuda_device)
    full = blp_b200.rank_sweep(model, e, rw, triples)
    h_rows, t_rows = e[triples[:, 0]], e[triples[:, 1]]
    acc_gt, acc_ge = torch.zeros_like(full["gt"]), torch.zeros_like(full["ge"])
    for rank in range(3):
        lo, hi = blp_b200.shard_bounds(n, 3, rank)
        part = blp_b200.rank_sweep(model, e[lo:hi].contiguous(), rw, triples, ent_offset=lo, h_rows=h_rows, t_rows=t_rows)
        assert torch.equal(part["true_score"], full["true_score"])
        acc_gt += part["gt"]
        acc_ge += part["ge"]
    assert torch.equal(acc_gt, full["gt"]) and torch.equal(acc_ge, full["ge"])


@pytest.mark.parametrize("d", (4, 32, 36, 64, 300, 384, 388, 768, 772, 1536, 2048))
@pytest.mark.parametrize("n,b", [(130, 1), (257, 3), (1000, 9), (3000, 40)])
def test_wide_transe_sweep_vs_oracle(d, n, b, cuda_device):
    """TransE at the BOW encoder widths (utils.py:11-19: 300 / 768) and every register-tile shape of the TMA-tiled
    wide kernel (blp_sweep_wide.cu): rows stream through shared memory in 64-float stages, zero padded to 32 floats."""
    ent, rel, heads, tails, rels = make_inputs("transe", n, d, b, seed=d + n)
    ent[n - 1] = ent[3]                # exact ties
    heads[0] = 3
    co = c_oracle.eval_rank("transe", ent.numpy(), ent[heads].numpy(), ent[tails].numpy(), rel[rels].numpy(),
                            heads.numpy(), tails.numpy())
    e, r = ent.to(cuda_device), rel.to(cuda_device)
    triples = torch.stack([heads, tails, rels], 1).to(cuda_device)
    out = blp_b200.rank_sweep("transe", e, r, triples)
    assert np.array_equal(out["true_score"].reshape(-1).cpu().numpy(), co["true_score"])
    assert np.array_equal(out["gt"].reshape(-1).cpu().numpy(), co["gt"])
    assert np.array_equal(out["ge"].reshape(-1).cpu().numpy(), co["ge"])
    # two row shards add up (entity-sharded sweeps at BOW widths)
    cut = n // 3
    h_rows, t_rows = e[triples[:, 0]], e[triples[:, 1]]
    a = blp_b200.rank_sweep("transe", e[:cut].contiguous(), r, triples, h_rows=h_rows, t_rows=t_rows)
    c = blp_b200.rank_sweep("transe", e[cut:].contiguous(), r, triples, ent_offset=cut, h_rows=h_rows, t_rows=t_rows)
    assert torch.equal(a["gt"] + c["gt"], out["gt"]) and torch.equal(a["ge"] + c["ge"], out["ge"])


@pytest.mark.parametrize("model", ("transe", "distmult", "complex", "simple"))
def test_topk_sweep_matches_the_oracle_scores(model, cuda_device):
    """blp_b200.topk_sweep: the k best candidates per query are the head of the stable descending sort of the oracle's
    full score rows (train.py:146-147 bits), rows as global ids; the true entity's rank from rank_sweep is consistent
    with its position in the list."""
    n, t, k = 700, 20, 10
    ent, rel, heads, tails, rels = make_inputs(model, n, 128, t, seed=17, n_rel=7)
    dev = cuda_device
    triples = torch.stack([heads, tails, rels], dim=1).to(dev)
    out = blp_b200.topk_sweep(model, ent.to(dev), rel.to(dev), triples, k=k, chunk=8)
    co = c_oracle.eval_rank(model, ent.numpy(), ent[heads].numpy(), ent[tails].numpy(), rel[rels].numpy(), heads.numpy(),
                            tails.numpy(), want_scores=True)
    scores = torch.from_numpy(np.asarray(co["scores"])).reshape(2, t, n)
    for role in range(2):
        sv, si = torch.sort(scores[role], dim=1, descending=True, stable=True)
        assert torch.equal(out["scores"][role].cpu(), sv[:, :k])
        assert torch.equal(out["index"][role].cpu(), si[:, :k])
    ranks = blp_b200.rank_sweep(model, ent.to(dev), rel.to(dev), triples)
    true = torch.stack([heads, tails]).to(dev)
    hit = out["index"] == true.unsqueeze(-1)
    in_list, pos = hit.any(-1), hit.to(torch.int64).argmax(-1)
    untied = (ranks["ge"] - ranks["gt"]) == 1                  # no other candidate shares the true score
    assert torch.equal(in_list[untied], (ranks["gt"] < k)[untied])          # in the list <=> fewer than k strictly better
    sel = untied & in_list
    assert torch.equal(pos[sel], ranks["gt"][sel].to(torch.int64))          # ... at position = number of better candidates
