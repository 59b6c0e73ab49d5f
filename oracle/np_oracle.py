"""NumPy restatement of the BLP hot path (second, independent oracle).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Every fp32 operation is
a separate NumPy float32 ufunc call, so each mul/add rounds exactly once, like
the reference's un-fused ATen ops.  Summation orders follow the reference's
CPU kernels (see blp_oracle.c header); reference lines are cited per function.
"""
import math

import numpy as np

f32 = np.float32


def _ceil_log2(x):
    if x <= 2:
        return 1
    return (x - 1).bit_length()


def aten_sum_lastdim(x):
    """torch.sum(x, dim=-1) for contiguous fp32 rows (ATen SumKernel.cpp order)."""
    x = np.asarray(x, f32)
    L = x.shape[-1]
    lead = x.shape[:-1]
    vec_size, size_ilp = L // 8, (L // 8) // 4
    levels = [np.zeros(lead + (4, 8), f32) for _ in range(4)]
    if size_ilp > 0:
        v = x[..., :size_ilp * 32].reshape(lead + (size_ilp, 4, 8))
        power = max(4, _ceil_log2(size_ilp) // 4)
        step, mask0 = 1 << power, (1 << power) - 1
        i = 0
        while i + step <= size_ilp:
            for _ in range(step):
                levels[0] = levels[0] + v[..., i, :, :]
                i += 1
            for j in range(1, 4):
                levels[j] = levels[j] + levels[j - 1]
                levels[j - 1] = np.zeros(lead + (4, 8), f32)
                if i & (mask0 << (j * power)):
                    break
        while i < size_ilp:
            levels[0] = levels[0] + v[..., i, :, :]
            i += 1
        for j in range(1, 4):
            levels[0] = levels[0] + levels[j]
    acc = levels[0]
    a0 = acc[..., 0, :]
    for i in range(size_ilp * 4, vec_size):
        a0 = a0 + x[..., i * 8:(i + 1) * 8]
    for k in range(1, 4):
        a0 = a0 + acc[..., k, :]
    fin = np.zeros(lead, f32)
    for k in range(vec_size * 8, L):
        fin = fin + x[..., k]
    for l in range(8):
        fin = fin + a0[..., l]
    return fin


def seq_sum_lastdim(x):
    """Strictly sequential fp32 sum (torch.norm(p=1) accumulation order)."""
    x = np.asarray(x, f32)
    s = np.zeros(x.shape[:-1], f32)
    for j in range(x.shape[-1]):
        s = s + x[..., j]
    return s


def l2_normalize_rows(x, eps=1e-12):
    """F.normalize(x, dim=-1) (models.py:40-41) with ATen's CPU vector-norm order: 8 lanes, one accumulator per
    lane over consecutive 8-element blocks (separate mul / add roundings, no FMA), lanes folded sequentially from
    lane 0, then the scalar tail; sqrt; x / max(norm, eps).  Verified bit-equal to torch 2.11 CPU (gen_golden.py)."""
    x = np.asarray(x, f32)
    d = x.shape[-1]
    full = d - d % 8
    acc = np.zeros(x.shape[:-1] + (8,), f32)
    for b in range(0, full, 8):
        blk = x[..., b:b + 8]
        acc = acc + blk * blk
    s = acc[..., 0].copy()
    for lane in range(1, 8):
        s = s + acc[..., lane]
    for j in range(full, d):
        s = s + x[..., j] * x[..., j]
    norm = np.sqrt(s).astype(f32)
    return (x / np.maximum(norm, f32(eps))[..., None]).astype(f32)


def transe_score(heads, tails, rels):
    """models.py:222-223."""
    heads, tails, rels = (np.asarray(a, f32) for a in (heads, tails, rels))
    return -seq_sum_lastdim(np.abs((heads + rels) - tails))


def distmult_score(heads, tails, rels):
    """models.py:226-227."""
    heads, tails, rels = (np.asarray(a, f32) for a in (heads, tails, rels))
    return aten_sum_lastdim(np.ascontiguousarray((heads * rels) * tails))


def _halves(x):
    L = x.shape[-1] // 2
    return x[..., :L], x[..., L:]


def complex_score(heads, tails, rels):
    """models.py:230-239."""
    heads, tails, rels = (np.asarray(a, f32) for a in (heads, tails, rels))
    hr, hi = _halves(heads)
    tr, ti = _halves(tails)
    rr, ri = _halves(rels)
    p = (((rr * hr) * tr + (rr * hi) * ti) + (ri * hr) * ti) - (ri * hi) * tr
    return aten_sum_lastdim(np.ascontiguousarray(p))


def simple_score(heads, tails, rels):
    """models.py:242-248."""
    heads, tails, rels = (np.asarray(a, f32) for a in (heads, tails, rels))
    hh, ht = _halves(heads)
    th, tt = _halves(tails)
    ra, rb = _halves(rels)
    p = (hh * ra) * tt + (th * rb) * ht
    return aten_sum_lastdim(np.ascontiguousarray(p)) / f32(2)


SCORE_FNS = {"transe": transe_score, "distmult": distmult_score,
             "complex": complex_score, "simple": simple_score}


def get_metrics(pred_scores, true_idx, k_values):
    """utils.py:86-111."""
    pred = np.asarray(pred_scores, f32)
    true_idx = np.asarray(true_idx).reshape(-1, 1)
    true = np.take_along_axis(pred, true_idx, axis=1)
    best = (pred > true).sum(1, keepdims=True) + 1
    worst = (pred >= true).sum(1, keepdims=True)
    avg = (best + worst).astype(f32) * f32(0.5)
    return f32(1) / avg, avg <= np.asarray(k_values).reshape(1, -1).astype(f32)


def margin_loss(pos, neg):
    """models.py:251-254 (mean accumulated in float64: tolerance quantity)."""
    m = (f32(1) - np.asarray(pos, f32)) + np.asarray(neg, f32)
    m = np.where(m < 0, f32(0), m)
    return f32(m.astype(np.float64).mean())


def _softplus(x):
    x = np.asarray(x, np.float64)
    return np.where(x > 20, x, np.log1p(np.exp(np.minimum(x, 20))))


def nll_loss(pos, neg):
    """models.py:257-258."""
    return f32((_softplus(-np.asarray(pos, np.float64)).mean() + _softplus(neg).mean()) / 2)


def l2_regularization(heads, tails, rels):
    """models.py:261-266."""
    return f32(sum((np.asarray(t, np.float64) ** 2).mean() for t in (heads, tails, rels)) / 3.0)


def compute_loss(model, loss, ent_embs, rel_rows, neg_idx, regularizer=0.0):
    """models.py:51-70 forward.  ent_embs (B,2,D), rel_rows (B,D), neg_idx (B,K,2)."""
    ent_embs = np.asarray(ent_embs, f32)
    B, _, D = ent_embs.shape
    rels = np.asarray(rel_rows, f32).reshape(B, 1, D)
    fn = SCORE_FNS[model]
    heads, tails = ent_embs[:, 0:1], ent_embs[:, 1:2]
    pos = fn(heads, tails, rels)
    neg_embs = ent_embs.reshape(2 * B, D)[np.asarray(neg_idx)]
    neg = fn(neg_embs[:, :, 0], neg_embs[:, :, 1], rels)
    ml = margin_loss(pos, neg) if loss == "margin" else nll_loss(pos, neg)
    reg = f32(regularizer) * l2_regularization(heads, tails, rels) if regularizer > 0 else f32(0)
    return f32(ml + reg), pos, neg


def eval_rank_batch(model, ent, heads, tails, rel_rows, k_values=(1, 3, 10), filter_mask=None):
    """train.py:141-171 for one batch; returns (pred (2B,N), gt, ge, recip, hits[, filtered...])."""
    ent = np.asarray(ent, f32)
    heads = np.asarray(heads).reshape(-1)
    tails = np.asarray(tails).reshape(-1)
    B = heads.shape[0]
    fn = SCORE_FNS[model]
    e = ent[None]
    h, t, r = ent[heads][:, None], ent[tails][:, None], np.asarray(rel_rows, f32)[:, None]
    pred = np.concatenate([fn(e, t, r), fn(h, e, r)])
    true = np.concatenate([heads, tails])
    out = {"pred": pred, "true": true}
    st = pred[np.arange(2 * B), true][:, None]
    out["gt"] = (pred > st).sum(1)
    out["ge"] = (pred >= st).sum(1)
    out["recip"], out["hits"] = get_metrics(pred, true, k_values)
    if filter_mask is not None:
        pf = pred.copy()
        pf[filter_mask] = pred.min() - f32(1.0)
        st = pf[np.arange(2 * B), true][:, None]
        out["gt_f"] = (pf > st).sum(1)
        out["ge_f"] = (pf >= st).sum(1)
        out["recip_f"], out["hits_f"] = get_metrics(pf, true, k_values)
    return out


# ---- tolerance reference for the tensor-core ("fast") mode -----------------------------------------
def folded_queries(model, h_rows, t_rows, r_rows):
    """Coefficient vectors of the bilinear scores with the candidate factored out, in fp64.

    score(candidate e) = C[q] . e  for head prediction (candidate plays `heads`, train.py:146) and tail
    prediction (candidate plays `tails`, train.py:147).  Derived term by term from models.py:226-248:
      distmult  sum (h r) t                                   -> head: r*t            tail: h*r
      complex   sum rr hr tr + rr hi ti + ri hr ti - ri hi tr -> head: (rr tr + ri ti | rr ti - ri tr)
                                                                 tail: (rr hr - ri hi | rr hi + ri hr)
      simple    1/2 sum hh ra tt + th rb ht                   -> head: (ra tt | th rb)/2
                                                                 tail: (rb ht | hh ra)/2
    Returns (C_head, C_tail), each (B, D) float64.  Used only to classify the fast mode's tolerance.
    """
    h, t, r = (np.asarray(x, dtype=np.float64) for x in (h_rows, t_rows, r_rows))
    if model == "distmult":
        return r * t, h * r
    L = h.shape[-1] // 2
    h1, h2, t1, t2, r1, r2 = h[:, :L], h[:, L:], t[:, :L], t[:, L:], r[:, :L], r[:, L:]
    if model == "complex":
        return (np.concatenate([r1 * t1 + r2 * t2, r1 * t2 - r2 * t1], axis=1),
                np.concatenate([r1 * h1 - r2 * h2, r1 * h2 + r2 * h1], axis=1))
    if model == "simple":
        return (np.concatenate([r1 * t2, t1 * r2], axis=1) * 0.5,
                np.concatenate([r2 * h2, h1 * r1], axis=1) * 0.5)
    raise ValueError(f"no contraction form for {model}")


def fast_mode_reference(model, ent, h_rows, t_rows, r_rows):
    """fp64 score matrix (2B, N) of the contraction form and its term mass sum_d |c_d e_d| (the tolerance scale)."""
    ch, ct = folded_queries(model, h_rows, t_rows, r_rows)
    c = np.concatenate([ch, ct], axis=0)
    e = np.asarray(ent, dtype=np.float64)
    return c @ e.T, np.abs(c) @ np.abs(e).T
