"""The CPU oracle against the golden vectors produced by the UNMODIFIED reference
(oracle/gen_golden.py).  Bit-exact for scores and integer ranks; 1e-5 for the
loss scalar and gradients (the reference's `mean` is not order-stable)."""
import numpy as np
import pytest

from conftest import golden, golden_names, mask_to_csr, unpack_mask
from oracle import c_oracle, np_oracle

EVAL = golden_names("eval_") 
EVAL = [n for n in EVAL if not n.startswith("eval_loop")]
TRAIN = golden_names("train_")


def _model(name):
    return name.split("_")[1]


@pytest.mark.parametrize("name", EVAL)
def test_c_oracle_eval_bits(name):
    g = golden(name)
    model = _model(name)
    heads, tails, rels = g["heads"][:, 0], g["tails"][:, 0], g["rels"][:, 0]
    mask = unpack_mask(g)
    indptr, idx = mask_to_csr(mask)
    out = c_oracle.eval_rank(model, g["ent"], g["ent"][heads], g["ent"][tails], g["rel"][rels], heads, tails,
                             indptr, idx, want_scores=True)
    assert np.array_equal(out["scores"], g["pred"])
    for k in ("gt", "ge", "gt_f", "ge_f"):
        assert np.array_equal(out[k], g[k]), k
    recip, hits = c_oracle.metrics_from_counts(out["gt"], out["ge"], [1, 3, 10])
    assert np.array_equal(recip, g["recip"]) and np.array_equal(hits, g["hits"])
    recip, hits = c_oracle.metrics_from_counts(out["gt_f"], out["ge_f"], [1, 3, 10])
    assert np.array_equal(recip, g["recip_f"]) and np.array_equal(hits, g["hits_f"])


@pytest.mark.parametrize("name", EVAL)
def test_np_oracle_eval_bits(name):
    g = golden(name)
    out = np_oracle.eval_rank_batch(_model(name), g["ent"], g["heads"], g["tails"], g["rel"][g["rels"][:, 0]],
                                    filter_mask=unpack_mask(g))
    assert np.array_equal(out["pred"], g["pred"])
    assert np.array_equal(out["gt"], g["gt"]) and np.array_equal(out["ge"], g["ge"])
    assert np.array_equal(out["gt_f"], g["gt_f"]) and np.array_equal(out["ge_f"], g["ge_f"])
    assert np.array_equal(out["recip_f"], g["recip_f"]) and np.array_equal(out["hits_f"], g["hits_f"])


@pytest.mark.parametrize("name", TRAIN)
def test_c_oracle_train(name):
    from conftest import strided_neg_idx
    g = golden(name)
    _, model, loss = name.split("_")
    rel_rows = g["rel_weight"][g["rels"][:, 0]]
    neg = strided_neg_idx(g).numpy()
    out = c_oracle.train_loss(model, loss, g["ent_embs"], rel_rows, neg, float(g["regularizer"]))
    assert np.array_equal(out["pos_scores"], g["pos_scores"][:, 0])
    assert np.array_equal(out["neg_scores"], g["neg_scores"])
    assert abs(float(out["loss"]) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    scale = np.abs(g["grad_ent"]).max()
    assert np.abs(out["grad_ent"] - g["grad_ent"]).max() <= 1e-5 * scale
    gw = np.zeros_like(g["grad_rel_weight"], dtype=np.float64)
    np.add.at(gw, g["rels"][:, 0], out["grad_rel"].astype(np.float64))
    assert np.abs(gw - g["grad_rel_weight"]).max() <= 1e-5 * max(1e-30, np.abs(g["grad_rel_weight"]).max())


@pytest.mark.parametrize("name", TRAIN)
def test_np_oracle_train(name):
    g = golden(name)
    _, model, loss = name.split("_")
    rel_rows = g["rel_weight"][g["rels"][:, 0]]
    l, pos, neg = np_oracle.compute_loss(model, loss, g["ent_embs"], rel_rows, g["neg_idx"], float(g["regularizer"]))
    assert np.array_equal(pos, g["pos_scores"]) and np.array_equal(neg, g["neg_scores"])
    assert abs(float(l) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))


def test_kat():
    g = golden("kat")
    assert abs(float(g["margin"]) - 0.4166667) < 1e-6
    assert abs(float(g["nll"]) - 0.6128115) < 1e-6
    assert np.allclose(g["margin_dneg"], [[1 / 6, 0, 1 / 6], [1 / 6, 1 / 6, 1 / 6]])
    assert np.allclose(g["margin_dpos"], [[-1 / 3], [-1 / 2]])
    assert g["ent2idx"].tolist() == [2, -1, -1, -1, 0, 1]
    assert tuple(g["neg_idx_shape"]) == (4, 3, 2) and tuple(g["neg_idx_stride"]) == (2, 8, 1)
    pos = np.array([[1.0], [1.0]], np.float32)
    neg = np.array([[0.0, -1.0, 2.0], [0.5, 0.0, 0.0]], np.float32)
    assert abs(float(np_oracle.margin_loss(pos, neg)) - float(g["margin"])) < 1e-6
    assert abs(float(np_oracle.nll_loss(np.float32([[0.3]]), np.float32([[0.1, -0.2]]))) - float(g["nll"])) < 1e-6
    s = np_oracle.transe_score(np.float32([[0, 1, -2]]), np.float32([[0, .5, 1]]), np.zeros((1, 3), np.float32))
    assert np.array_equal(s, g["transe"])


def test_aten_sum_order_self_consistent():
    """C and NumPy restatements of ATen's row sum agree bit for bit on widths beyond the fixtures."""
    rng = np.random.default_rng(0)
    for L in (8, 63, 64, 100, 128, 150, 300, 384, 768, 800, 2048):
        x = rng.standard_normal((7, L)).astype(np.float32)
        a = np_oracle.aten_sum_lastdim(x)
        b = np.array([c_oracle.aten_sum(v) for v in x])
        assert np.array_equal(a, b), L


def test_np_oracle_normalize_bits():
    """F.normalize through the reference's encode() (models.py:38-43): ATen vector-norm order, any width."""
    g = golden("normalize")
    for d in (128, 300, 768, 100):
        assert np.array_equal(np_oracle.l2_normalize_rows(g[f"x_{d}"]), g[f"y_{d}"]), d


@pytest.mark.parametrize("model", ("transe", "distmult", "complex", "simple"))
def test_torch_port_equals_reference(model):
    """oracle/torch_port.py (the CPU arm's fallback when oracle/_ref is absent) against the reference's OWN functions
    (imported from /root/reference, or byte-compiled under oracle/_ref): compute_loss forward + backward with the
    reference sampler's strided neg_idx (models.py:51-70), and the eval statements train.py:141-153 / 164-167 --
    bit for bit (single-threaded, so the mean reductions are order-stable too)."""
    import torch
    from oracle import ref_loader, torch_port
    if ref_loader.available() is None:
        pytest.skip("reference modules not available (python oracle/build_ref.py where /root/reference is mounted)")
    mods = ref_loader.load(("models", "utils", "data"))
    ref_models, ref_utils, ref_data = mods["models"], mods["utils"], mods["data"]
    threads = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        torch.manual_seed(3)
        b, k, d, n, n_rel = 16, 24, 128, 300, 7
        for loss in ("margin", "nll"):
            reg = 1e-3 if model == "complex" else 0
            m = ref_models.TransductiveLinkPrediction(d, model, loss, n, n_rel, reg)
            pairs = torch.randint(0, n, (b, 2))
            rels = torch.randint(0, n_rel, (b, 1))
            neg_idx = ref_data.get_negative_sampling_indices(b, k)            # strided view, data.py:78-79
            x_ref = m.encode(pairs).detach().clone().requires_grad_(True)
            want = m.compute_loss(x_ref, rels, neg_idx)
            want.backward()
            x = x_ref.detach().clone().requires_grad_(True)
            w = m.rel_emb.weight.detach().clone().requires_grad_(True)
            got = torch_port.batch_loss(model, loss, x, w[rels[:, 0]], neg_idx, reg)
            got.backward()
            assert torch.equal(got.detach(), want.detach()), (model, loss)
            assert torch.equal(x.grad, x_ref.grad) and torch.equal(w.grad, m.rel_emb.weight.grad)
        # eval statements, raw and filtered
        ent_emb = m.encode(torch.arange(n)).detach().unsqueeze(0)
        heads, tails = torch.randint(0, n, (b, 1)), torch.randint(0, n, (b, 1))
        rel_embs = m.rel_emb(rels).detach()
        k_values = torch.tensor([[1, 3, 10]])
        with torch.no_grad():
            head_embs, tail_embs = ent_emb.squeeze()[heads], ent_emb.squeeze()[tails]
            pred = torch.cat((m.score_fn(ent_emb, tail_embs, rel_embs), m.score_fn(head_embs, ent_emb, rel_embs)))
            true = torch.cat((heads, tails))
            recip, hits = ref_utils.get_metrics(pred, true, k_values)
            mask = torch.rand(2 * b, n) < 0.03
            mask[torch.arange(2 * b), true[:, 0]] = False
            pred_f = pred.clone()
            pred_f[mask] = pred_f.min() - 1.0
            recip_f, hits_f = ref_utils.get_metrics(pred_f, true, k_values)
        out = torch_port.eval_batch(model, ent_emb, heads, tails, rel_embs, k_values, filter_mask=mask)
        assert torch.equal(out["recip"], recip) and torch.equal(out["hits"], hits)
        assert torch.equal(out["recip_f"], recip_f) and torch.equal(out["hits_f"], hits_f)
        assert torch.equal(out["pred"], pred_f)
    finally:
        torch.set_num_threads(threads)
