"""BASELINE.json configs at their full sizes, through size-independent properties plus oracle checks on the
queries the oracle can finish in seconds:
  configs[3]  WN18RR-sized (40,943 entities) BLP-ComplEx full-entity sweep, exact and tensor-core mode
  configs[4]  one 8-way shard of the Wikidata5M-scale table (600,000 rows) BLP-TransE at the reference's
              eval batch of 2 (the HBM-bound register-tile shape), with ent_offset / shard sums
  train       B = 1024, K = 512 (the reference's Wikidata5M training batch) against the oracle
and the degenerate shapes (T = 0, K = 0)."""
import numpy as np
import pytest
import torch

from oracle import c_oracle

import blp_b200
from blp_b200 import ops
from test_gpu_eval import make_inputs

pytestmark = pytest.mark.gpu


def test_wn18rr_complex_full_entity_sweep(cuda_device):
    model, n, b = "complex", 40943, 64                                  # reference eval_batch_size = 64
    ent, rel, heads, tails, rels = make_inputs(model, n, 128, b, seed=17, n_rel=11)
    dev = cuda_device
    e, r = ent.to(dev), rel.to(dev)
    triples = torch.stack([heads, tails, rels], dim=1).to(dev)
    full = blp_b200.rank_sweep(model, e, r, triples)
    assert bool((full["gt"] < full["ge"]).all())                        # the true entity ties itself
    # oracle on every 8th triple (16 queries x 40,943 candidates)
    sel = torch.arange(0, b, 8)
    co = c_oracle.eval_rank(model, ent.numpy(), ent[heads[sel]].numpy(), ent[tails[sel]].numpy(), rel[rels[sel]].numpy(),
                            heads[sel].numpy(), tails[sel].numpy())
    k = len(sel)
    assert np.array_equal(full["gt"].cpu()[:, sel].reshape(-1).numpy(), co["gt"])
    assert np.array_equal(full["ge"].cpu()[:, sel].reshape(-1).numpy(), co["ge"])
    assert np.array_equal(full["true_score"].cpu()[:, sel].reshape(-1).numpy(), co["true_score"]) and k == 8
    # candidate order does not matter; shard sums are exact
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(2)).to(dev)
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(n, device=dev)
    tp = torch.stack([inv[triples[:, 0]], inv[triples[:, 1]], triples[:, 2]], dim=1)
    shuffled = blp_b200.rank_sweep(model, e[perm].contiguous(), r, tp)
    assert torch.equal(shuffled["gt"], full["gt"]) and torch.equal(shuffled["ge"], full["ge"])
    acc = None
    for s in range(3):
        lo, hi = blp_b200.shard_bounds(n, 3, s)
        part = blp_b200.rank_sweep(model, e[lo:hi].contiguous(), r, triples, ent_offset=lo,
                                   h_rows=e[triples[:, 0]], t_rows=e[triples[:, 1]])
        cur = torch.stack([part["gt"], part["ge"]])
        acc = cur if acc is None else acc + cur
    assert torch.equal(acc[0], full["gt"]) and torch.equal(acc[1], full["ge"])
    # tensor-core mode: same true scores, ranks equal except inside the tolerance band
    fast = blp_b200.rank_sweep(model, e, r, triples, mode="fast")
    assert torch.equal(fast["true_score"], full["true_score"])
    assert ((fast["gt"] == full["gt"]) & (fast["ge"] == full["ge"])).float().mean().item() >= 0.95
    assert int((fast["gt"] - full["gt"]).abs().max()) <= 3
    assert abs(blp_b200.finalize(fast)["mrr"] - blp_b200.finalize(full)["mrr"]) <= 1e-6


def test_wikidata5m_shard_transe_eval_batch_2(cuda_device):
    model, n_shard, b, offset = "transe", 600_000, 2, 1_200_000          # rank 2 of 8: rows [1.2 M, 1.8 M)
    g = torch.Generator().manual_seed(23)
    shard = torch.nn.functional.normalize(torch.randn(n_shard, 128, generator=g), dim=-1)
    rel = (torch.rand(822, 128, generator=g) * 2 - 1) * (6.0 / (822 + 128)) ** 0.5
    # one true head lives in this shard, the other rows come from other ranks (passed pre-gathered)
    h_rows = torch.stack([shard[12345], torch.nn.functional.normalize(torch.randn(128, generator=g), dim=-1)])
    t_rows = torch.nn.functional.normalize(torch.randn(2, 128, generator=g), dim=-1)
    triples = torch.tensor([[offset + 12345, 7, 3], [4_000_000, 99, 800]])
    dev = cuda_device
    out = blp_b200.rank_sweep(model, shard.to(dev), rel.to(dev), triples.to(dev), ent_offset=offset,
                              h_rows=h_rows.to(dev), t_rows=t_rows.to(dev), chunk=2)
    co = c_oracle.eval_rank(model, shard.numpy(), h_rows.numpy(), t_rows.numpy(), rel[triples[:, 2]].numpy(),
                            None, None)                   # true rows = the query rows themselves (sharded caller)
    assert np.array_equal(out["true_score"].reshape(-1).cpu().numpy(), co["true_score"])
    assert np.array_equal(out["gt"].reshape(-1).cpu().numpy(), co["gt"])
    assert np.array_equal(out["ge"].reshape(-1).cpu().numpy(), co["ge"])
    # the head query of triple 0 scans its own true row in this shard: ge counts the self-match, gt does not
    assert int(out["ge"][0, 0]) >= int(out["gt"][0, 0]) + 1
    # sub-shards add up
    a = blp_b200.rank_sweep(model, shard[:250_001].contiguous().to(dev), rel.to(dev), triples.to(dev), ent_offset=offset,
                            h_rows=h_rows.to(dev), t_rows=t_rows.to(dev))
    c = blp_b200.rank_sweep(model, shard[250_001:].contiguous().to(dev), rel.to(dev), triples.to(dev),
                            ent_offset=offset + 250_001, h_rows=h_rows.to(dev), t_rows=t_rows.to(dev))
    assert torch.equal(a["gt"] + c["gt"], out["gt"]) and torch.equal(a["ge"] + c["ge"], out["ge"])


@pytest.mark.parametrize("model,loss", [("transe", "margin"), ("distmult", "nll")])
def test_train_wikidata5m_batch(model, loss, cuda_device):
    """B = 1024 positives x K = 512 in-batch negatives with the device sampler's strided indices."""
    b, k, d, n_rel = 1024, 512, 128, 822
    g = torch.Generator().manual_seed(31)
    ent = torch.randn(b, 2, d, generator=g)
    if model == "transe":
        ent = torch.nn.functional.normalize(ent, dim=-1)
    rel_w = (torch.rand(n_rel, d, generator=g) * 2 - 1) * (6.0 / (n_rel + d)) ** 0.5
    rels = torch.randint(0, n_rel, (b, 1), generator=g)
    neg = blp_b200.get_negative_sampling_indices(b, k, device=cuda_device, seed=77)
    res = ops.train_loss(model, loss, ent.to(cuda_device), rel_w.to(cuda_device), rels.to(cuda_device), neg,
                         want_grad=True)
    co = c_oracle.train_loss(model, loss, ent.numpy(), rel_w[rels[:, 0]].numpy(), neg.cpu().numpy(), 0.0)
    assert abs(res["loss"].item() - float(co["loss"])) <= 1e-5 * abs(float(co["loss"]))
    ge = res["grad_ent"].cpu().numpy()
    assert np.abs(ge - co["grad_ent"]).max() <= 2e-5 * np.abs(co["grad_ent"]).max()
    assert not ops.index_error_flag(res)


def test_degenerate_shapes(cuda_device):
    dev = cuda_device
    ent, rel = torch.randn(50, 128, device=dev), torch.randn(3, 128, device=dev)
    empty = torch.empty((0, 3), dtype=torch.int64, device=dev)
    out = blp_b200.rank_sweep("distmult", ent, rel, empty)
    assert out["gt"].shape == (2, 0) and out["sums"].tolist() == [0.0, 0.0, 0.0, 0.0]
    didx = blp_b200.DeviceFilterIndex(np.array([[1, 2, 0]]), None, 50, 3, dev)
    out = blp_b200.rank_sweep("distmult", ent, rel, empty, filter_index=didx)
    assert out["gt_f"].shape == (2, 0)
    neg = blp_b200.get_negative_sampling_indices(4, 0, device=dev)
    assert tuple(neg.shape) == (4, 0, 2)
    one = torch.tensor([[3, 3, 1]], device=dev)                          # head == tail, single triple
    out = blp_b200.rank_sweep("simple", ent, rel, one)
    assert out["gt"].shape == (2, 1) and bool((out["gt"] < out["ge"]).all())


def test_relation_sorted_sweep_is_bit_identical(cuda_device):
    """rank_sweep(sort_by_relation=True) runs the triples in relation order so the TransE kernel can share
    fl(candidate + r) across head queries; outputs come back in the caller's order with the same bits."""
    model, n, t = "transe", 14541, 4608                                  # 67 M scores per direction: above the threshold
    ent, rel, heads, tails, rels = make_inputs(model, n, 128, t, seed=41, n_rel=37)
    dev = cuda_device
    e, r = ent.to(dev), rel.to(dev)
    triples = torch.stack([heads, tails, rels], dim=1).to(dev)
    a = blp_b200.rank_sweep(model, e, r, triples, sort_by_relation=True)
    b = blp_b200.rank_sweep(model, e, r, triples, sort_by_relation=False)
    for k in ("gt", "ge", "true_score", "recip", "hits"):
        assert torch.equal(a[k], b[k]), k
    # a chunk that is already relation-sorted goes through the shared-relation code path of the kernel
    order = torch.argsort(triples[:, 2], stable=True)
    c = blp_b200.rank_sweep(model, e, r, triples[order].contiguous(), sort_by_relation=False)
    assert torch.equal(c["gt"], b["gt"][:, order]) and torch.equal(c["ge"], b["ge"][:, order])
    sel = order[:4].cpu()
    co = c_oracle.eval_rank(model, ent.numpy(), ent[heads[sel]].numpy(), ent[tails[sel]].numpy(), rel[rels[sel]].numpy(),
                            heads[sel].numpy(), tails[sel].numpy())
    assert np.array_equal(c["gt"].cpu()[:, :4].reshape(-1).numpy(), co["gt"])
    assert np.array_equal(c["ge"].cpu()[:, :4].reshape(-1).numpy(), co["ge"])


@pytest.mark.parametrize("d", (128, 768))
def test_umls_plumbing_config(d, cuda_device):
    """BASELINE configs[0]: UMLS-sized (135 entities, 46 relations) TransE, margin loss, regularizer 1e-2, B = 64,
    K = 64 -- at dim 128 (BASELINE.json) and at the 768 of the bert-bow model scripts/test-umls.sh actually runs.
    Everything is small enough for the oracle to check every number."""
    model, n, n_rel, b, k = "transe", 135, 46, 64, 64
    ent, rel, heads, tails, rels = make_inputs(model, n, d, b, seed=d, n_rel=n_rel)
    dev = cuda_device
    # eval: one batch of 64 test triples against all 135 entities
    out = blp_b200.rank_sweep(model, ent.to(dev), rel.to(dev), torch.stack([heads, tails, rels], dim=1).to(dev))
    co = c_oracle.eval_rank(model, ent.numpy(), ent[heads].numpy(), ent[tails].numpy(), rel[rels].numpy(),
                            heads.numpy(), tails.numpy())
    for key in ("gt", "ge", "true_score"):
        assert np.array_equal(out[key].reshape(-1).cpu().numpy(), co[key]), key
    recip, hits = c_oracle.metrics_from_counts(co["gt"], co["ge"], [1, 3, 10])
    assert np.array_equal(out["recip"].reshape(-1).cpu().numpy(), recip.reshape(-1))
    assert np.array_equal(out["hits"].cpu().numpy(), hits)
    # train: compute_loss forward + backward with the device sampler's indices
    g = torch.Generator().manual_seed(3)
    pairs = torch.randint(0, n, (b, 2), generator=g)
    brels = torch.randint(0, n_rel, (b, 1), generator=g)
    neg = blp_b200.get_negative_sampling_indices(b, k, device=dev, seed=11)
    m = blp_b200.TransductiveLinkPrediction(d, model, "margin", n, n_rel, 1e-2).to(dev)
    with torch.no_grad():
        m.ent_emb.weight.copy_(ent)
        m.rel_emb.weight.copy_(rel)
    ent_embs = m.encode(pairs.to(dev)).detach().requires_grad_(True)
    loss = m.compute_loss(ent_embs, brels.to(dev), neg)
    loss.backward()
    co = c_oracle.train_loss(model, "margin", ent_embs.detach().cpu().numpy(), rel[brels[:, 0]].numpy(), neg.cpu().numpy(), 1e-2)
    assert abs(loss.item() - float(co["loss"])) <= 1e-5 * abs(float(co["loss"]))
    assert np.abs(ent_embs.grad.cpu().numpy() - co["grad_ent"]).max() <= 2e-5 * np.abs(co["grad_ent"]).max()


@pytest.mark.parametrize("model", ("transe", "distmult", "complex", "simple"))
@pytest.mark.parametrize("n_rel", (3, 37, 500))
def test_aligned_triples_sweep_is_bit_identical(model, n_rel, cuda_device):
    """blp_b200.AlignedTriples: relation-sorted order with every relation's run padded to a multiple of 4, so every warp of
    the TransE kernel shares fl(candidate + r) across its head queries.  Outputs come back in the caller's order with the
    same bits as the unsorted sweep; padding entries are dropped."""
    n, t = 3000, 1500
    ent, rel, heads, tails, rels = make_inputs(model, n, 128, t, seed=n_rel, n_rel=n_rel)
    dev = cuda_device
    e, r = ent.to(dev), rel.to(dev)
    triples = torch.stack([heads, tails, rels], dim=1).to(dev)
    aligned = blp_b200.AlignedTriples(triples)
    assert aligned.num_padded % 4 == 0 and aligned.num_padded >= t
    pr = aligned.padded[:, 2].reshape(-1, 4)
    assert bool((pr == pr[:, :1]).all()) and torch.equal(aligned.padded[aligned.dest], triples)
    a = blp_b200.rank_sweep(model, e, r, aligned)
    b = blp_b200.rank_sweep(model, e, r, triples, sort_by_relation=False)
    for k in ("gt", "ge", "true_score", "recip", "hits"):
        assert torch.equal(a[k], b[k]), k
    assert abs(float(a["sums"][0]) - float(b["sums"][0])) < 1e-9
    # with a device filter index, and with pre-gathered rows on a shard
    edges = np.stack([heads.numpy()[:400], np.roll(tails.numpy()[:400], 1), rels.numpy()[:400]], 1)
    fidx = blp_b200.DeviceFilterIndex(edges, None, n, n_rel, dev)
    af = blp_b200.rank_sweep(model, e, r, aligned, filter_index=fidx)
    bf = blp_b200.rank_sweep(model, e, r, triples, filter_index=fidx, sort_by_relation=False)
    assert torch.equal(af["gt_f"], bf["gt_f"]) and torch.equal(af["ge_f"], bf["ge_f"])
    h_rows, t_rows = e[triples[:, 0]], e[triples[:, 1]]
    lo = blp_b200.rank_sweep(model, e[:1000].contiguous(), r, aligned, h_rows=h_rows, t_rows=t_rows)
    hi = blp_b200.rank_sweep(model, e[1000:].contiguous(), r, aligned, ent_offset=1000, h_rows=h_rows, t_rows=t_rows)
    assert torch.equal(lo["gt"] + hi["gt"], b["gt"]) and torch.equal(lo["ge"] + hi["ge"], b["ge"])


@pytest.mark.gpu
@pytest.mark.parametrize("q,k_values", [(40960, [1, 3, 10]), (40963, [1, 3, 10]), (8192, [1, 3, 10, 100]), (40961, [1, 10]),
                                       (4095, [1, 3, 10])])
def test_metrics_of_a_whole_evaluation_set(q, k_values, cuda_device):
    """utils.py:106-109 + train.py:154-157 over a whole evaluation set in one launch (blp_rank_metrics): from 4,096
    queries on the kernel takes four queries per thread with 16-byte accesses and packed hit bytes (3 or 4 hit
    positions); reciprocal ranks, hits and the hit counts must equal the oracle's bit for bit, the MRR sum to 1e-12."""
    g = torch.Generator().manual_seed(q)
    gt = torch.randint(0, 14541, (q,), generator=g, dtype=torch.int32)
    gt[::7] = 0                                               # rank-1 queries: hits at every k
    ge = gt + torch.randint(1, 5, (q,), generator=g, dtype=torch.int32)
    recip, hits, sums = ops.rank_metrics(gt.to(cuda_device), ge.to(cuda_device), k_values)
    want_recip, want_hits = c_oracle.metrics_from_counts(gt.numpy(), ge.numpy(), k_values)
    assert np.array_equal(recip.cpu().numpy(), want_recip)
    assert np.array_equal(hits.cpu().numpy().astype(np.uint8), want_hits.astype(np.uint8))
    s = sums.cpu().numpy()
    assert np.array_equal(s[1:], want_hits.astype(np.float64).sum(0))
    ref = want_recip.astype(np.float64).sum()
    assert abs(s[0] - ref) <= 1e-12 * ref
