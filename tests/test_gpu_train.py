"""CUDA fused compute_loss / loss functions vs the golden vectors (reference autograd) and the oracle.
Scores, loss and gradients within 1e-5 relative (north_star); KAT conventions exact."""
import numpy as np
import pytest
import torch

from conftest import golden, golden_names, strided_neg_idx
from oracle import c_oracle

import blp_b200
from blp_b200 import ops

pytestmark = pytest.mark.gpu
TRAIN = golden_names("train_")
MODELS = ("transe", "distmult", "complex", "simple")
RTOL = 1e-5


def _close(a, b, scale=None, tol=RTOL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    scale = np.abs(b).max() if scale is None else scale
    return np.abs(a - b).max() <= tol * max(scale, 1e-30)


def _mass(model, ent_embs, rel_rows, neg_idx):
    """Sum of |terms| per score: the scale a 1e-5 relative bound on a cancelling sum refers to."""
    e = np.abs(ent_embs).astype(np.float64)
    return float(e.max() ** 2 * np.abs(rel_rows).max() * ent_embs.shape[-1]) if model != "transe" else float(ent_embs.shape[-1])


@pytest.mark.parametrize("name", TRAIN)
def test_golden_compute_loss(name, cuda_device):
    g = golden(name)
    _, model, loss = name.split("_")
    b = g["ent_embs"].shape[0]
    m = blp_b200.TransductiveLinkPrediction(128, model, loss, num_entities=4, num_relations=g["rel_weight"].shape[0],
                                            regularizer=float(g["regularizer"])).to(cuda_device)
    with torch.no_grad():
        m.rel_emb.weight.copy_(torch.from_numpy(g["rel_weight"]))
    ent_embs = torch.from_numpy(g["ent_embs"]).to(cuda_device).requires_grad_(True)
    neg_idx = strided_neg_idx(g)
    assert not neg_idx.is_contiguous()
    # move the strided view to the device preserving its strides (the sampler output goes through .to(device))
    base = torch.empty(neg_idx.untyped_storage().size() // 8, dtype=torch.int64, device=cuda_device)
    neg_dev = base.as_strided(neg_idx.shape, neg_idx.stride())
    neg_dev.copy_(neg_idx)
    out = m.compute_loss(ent_embs, torch.from_numpy(g["rels"]).to(cuda_device), neg_dev)
    assert out.dim() == 0
    out.backward()
    assert abs(out.item() - float(g["loss"])) <= RTOL * abs(float(g["loss"]))
    assert _close(ent_embs.grad.cpu().numpy(), g["grad_ent"])
    assert _close(m.rel_emb.weight.grad.cpu().numpy(), g["grad_rel_weight"])
    res = ops.train_loss(model, loss, ent_embs.detach(), m.rel_emb.weight.detach(), torch.from_numpy(g["rels"]).to(cuda_device),
                         neg_dev, regularizer=float(g["regularizer"]), want_grad=False, want_neg_scores=True)
    mass = _mass(model, g["ent_embs"], g["rel_weight"], g["neg_idx"])
    assert _close(res["pos_scores"].cpu().numpy(), g["pos_scores"][:, 0], scale=mass)
    assert _close(res["neg_scores"].cpu().numpy(), g["neg_scores"], scale=mass)
    assert not ops.index_error_flag(res)


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("loss", ("margin", "nll"))
@pytest.mark.parametrize("b,k,d", [(64, 512, 128), (5, 259, 128), (2, 2, 128), (33, 70, 128), (16, 64, 256), (8, 32, 768)])
def test_compute_loss_vs_oracle(model, loss, b, k, d, cuda_device):
    g = torch.Generator().manual_seed(b * 1000 + k)
    ent = torch.randn(b, 2, d, generator=g)
    if model == "transe":
        ent = torch.nn.functional.normalize(ent, dim=-1)
    n_rel = 7
    rel_w = (torch.rand(n_rel, d, generator=g) * 2 - 1) * (6.0 / (n_rel + d)) ** 0.5
    rels = torch.randint(0, n_rel, (b, 1), generator=g)
    neg = torch.randint(0, 2 * b, (b, k, 2), generator=g)
    reg = 1e-2 if model == "complex" else 0.0
    co = c_oracle.train_loss(model, loss, ent.numpy(), rel_w[rels[:, 0]].numpy(), neg.numpy(), reg)
    res = ops.train_loss(model, loss, ent.to(cuda_device), rel_w.to(cuda_device), rels.to(cuda_device), neg.to(cuda_device),
                         regularizer=reg, want_grad=True, want_neg_scores=True)
    assert abs(res["loss"].item() - float(co["loss"])) <= RTOL * abs(float(co["loss"]))
    mass = _mass(model, ent.numpy(), rel_w.numpy(), None)
    assert _close(res["pos_scores"].cpu().numpy(), co["pos_scores"], scale=mass)
    assert _close(res["neg_scores"].cpu().numpy(), co["neg_scores"], scale=mass)
    assert _close(res["grad_ent"].cpu().numpy(), co["grad_ent"], tol=2e-5)
    gw = np.zeros((n_rel, d), np.float64)
    np.add.at(gw, rels[:, 0].numpy(), co["grad_rel"].astype(np.float64))
    assert _close(res["grad_rel_weight"].cpu().numpy(), gw, tol=2e-5)


def test_kat_loss_conventions(cuda_device):
    """SURVEY.md Appendix B, regenerated from the reference (kat.npz)."""
    g = golden("kat")
    pos = torch.tensor([[1.0], [1.0]], device=cuda_device, requires_grad=True)
    neg = torch.tensor([[0.0, -1.0, 2.0], [0.5, 0.0, 0.0]], device=cuda_device, requires_grad=True)
    ml = blp_b200.margin_loss(pos, neg)
    ml.backward()
    assert abs(ml.item() - float(g["margin"])) < 1e-6
    assert np.allclose(pos.grad.cpu().numpy(), g["margin_dpos"], atol=1e-7)
    assert np.allclose(neg.grad.cpu().numpy(), g["margin_dneg"], atol=1e-7)     # loss == 0 keeps its gradient
    p2 = torch.tensor([[0.3]], device=cuda_device, requires_grad=True)
    n2 = torch.tensor([[0.1, -0.2]], device=cuda_device, requires_grad=True)
    nl = blp_b200.nll_loss(p2, n2)
    nl.backward()
    assert abs(nl.item() - float(g["nll"])) < 1e-6
    assert np.allclose(p2.grad.cpu().numpy(), g["nll_dpos"], atol=1e-7)
    assert np.allclose(n2.grad.cpu().numpy(), g["nll_dneg"], atol=1e-7)
    h = torch.tensor([[0.0, 1.0, -2.0]], device=cuda_device)
    t = torch.tensor([[0.0, 0.5, 1.0]], device=cuda_device)
    r = torch.zeros(1, 3, device=cuda_device)
    assert np.array_equal(blp_b200.transe_score(h, t, r).cpu().numpy(), g["transe"])


def test_l2_regularization(cuda_device):
    g = torch.Generator().manual_seed(0)
    h, t, r = (torch.randn(9, 1, 128, generator=g) for _ in range(3))
    want = sum((x.double() ** 2).mean() for x in (h, t, r)) / 3.0
    got = blp_b200.l2_regularization(h.to(cuda_device), t.to(cuda_device), r.to(cuda_device))
    assert abs(got.item() - want.item()) <= 1e-6 * want.item()


def test_out_of_range_index_is_flagged(cuda_device):
    b, k, d = 4, 8, 128
    ent = torch.randn(b, 2, d, device=cuda_device)
    rel = torch.randn(3, d, device=cuda_device)
    rels = torch.zeros(b, 1, dtype=torch.long, device=cuda_device)
    neg = torch.randint(0, 2 * b, (b, k, 2), device=cuda_device)
    neg[1, 2, 0] = 2 * b          # models.py:65 would raise an index error
    res = ops.train_loss("transe", "margin", ent, rel, rels, neg, want_grad=False)
    assert ops.index_error_flag(res)


def test_backward_scales_with_upstream_gradient(cuda_device):
    """train.py:344 takes .mean() over DataParallel replicas: backward must honour grad_output."""
    b, k, d = 8, 16, 128
    g = torch.Generator().manual_seed(3)
    ent = torch.randn(b, 2, d, generator=g).to(cuda_device).requires_grad_(True)
    m = blp_b200.TransductiveLinkPrediction(d, "distmult", "margin", 4, 5, 0).to(cuda_device)
    rels = torch.randint(0, 5, (b, 1), generator=g).to(cuda_device)
    neg = torch.randint(0, 2 * b, (b, k, 2), generator=g).to(cuda_device)
    loss = m.compute_loss(ent, rels, neg)
    (g1,) = torch.autograd.grad(loss, ent, retain_graph=True)
    (g3,) = torch.autograd.grad(loss * 3.0, ent)
    assert torch.allclose(g3, 3.0 * g1, rtol=1e-6, atol=0)


def test_reference_sampler_structure_full_size(cuda_device):
    """B=64, K=512 (BASELINE config 2) with a sampler-shaped neg_idx: one column is always the row's own entity."""
    b, k, d = 64, 512, 128
    g = torch.Generator().manual_seed(11)
    ent = torch.nn.functional.normalize(torch.randn(b, 2, d, generator=g), dim=-1)
    rel_w = (torch.rand(237, d, generator=g) * 2 - 1) * (6.0 / (237 + d)) ** 0.5
    rels = torch.randint(0, 237, (b, 1), generator=g)
    own = torch.arange(2 * b).reshape(b, 2)
    neg = own[:, None, :].repeat(1, k, 1)
    side = torch.randint(0, 2, (b, k), generator=g)
    repl = torch.randint(0, 2 * b, (b, k), generator=g)
    neg[torch.arange(b)[:, None], torch.arange(k)[None, :], side] = repl
    neg_t = neg.permute(1, 0, 2).contiguous().permute(1, 0, 2)          # non-contiguous like data.py:78-79
    assert not neg_t.is_contiguous()
    co = c_oracle.train_loss("transe", "margin", ent.numpy(), rel_w[rels[:, 0]].numpy(), neg.numpy(), 0.0)
    base = torch.empty(b * k * 2, dtype=torch.int64, device=cuda_device)
    neg_dev = base.as_strided(neg_t.shape, neg_t.stride())
    neg_dev.copy_(neg_t)
    res = ops.train_loss("transe", "margin", ent.to(cuda_device), rel_w.to(cuda_device), rels.to(cuda_device), neg_dev,
                         want_grad=True, want_neg_scores=True)
    assert abs(res["loss"].item() - float(co["loss"])) <= RTOL * abs(float(co["loss"]))
    assert _close(res["neg_scores"].cpu().numpy(), co["neg_scores"], scale=128.0)
    assert _close(res["grad_ent"].cpu().numpy(), co["grad_ent"], tol=2e-5)
    assert res["launches"] == 1


@pytest.mark.parametrize("model,loss", [("transe", "margin"), ("complex", "nll")])
def test_graphed_loss_step_equals_eager(model, loss, cuda_device):
    """blp_b200.GraphedLossStep replays exactly `compute_loss(...); backward()` (train.py:344-347): same loss and
    gradients as the eager path, for fresh inputs on every call, with the reference sampler's strided neg_idx."""
    b, k, d, n_rel = 32, 96, 128, 7
    g = torch.Generator().manual_seed(5)
    m = blp_b200.TransductiveLinkPrediction(d, model, loss, 100, n_rel, 1e-3 if model == "complex" else 0).to(cuda_device)
    step = blp_b200.GraphedLossStep(m, b, k)
    for it in range(3):
        x = torch.randn(b, 2, d, generator=g).to(cuda_device)
        rels = torch.randint(0, n_rel, (b, 1), generator=g).to(cuda_device)
        neg = blp_b200.get_negative_sampling_indices(b, k, device=cuda_device, seed=3, offset=it)
        loss_g, grad_g = step(x, rels, neg)
        loss_g, grad_g, grad_rel_g = loss_g.clone(), grad_g.clone(), m.rel_emb.weight.grad.clone()
        xe = x.clone().requires_grad_(True)
        m.rel_emb.weight.grad = None
        loss_e = m.compute_loss(xe, rels, neg)
        loss_e.backward()
        assert torch.equal(loss_g, loss_e.detach())
        # red.global.add arrival order differs from launch to launch: fp32 sums agree to a few ulp
        assert torch.allclose(grad_g, xe.grad, rtol=1e-5, atol=1e-7)
        assert torch.allclose(grad_rel_g, m.rel_emb.weight.grad, rtol=1e-5, atol=1e-7)
