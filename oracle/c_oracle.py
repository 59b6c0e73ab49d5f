"""ctypes binding of oracle/blp_oracle.c (numpy in, numpy out).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libblp_oracle.so")
_lib = None

MODELS = {"transe": 0, "distmult": 1, "complex": 2, "simple": 3}
LOSSES = {"margin": 0, "nll": 1}

_f32p = ctypes.POINTER(ctypes.c_float)
_i64p = ctypes.POINTER(ctypes.c_int64)
_u8p = ctypes.POINTER(ctypes.c_uint8)


def build(force=False):
    """Compile blp_oracle.c with the committed Makefile (system gcc, OpenMP)."""
    src = os.path.join(_HERE, "blp_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s", "libblp_oracle.so"] + (["-B"] if force else []), check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.blp_oracle_score_one.restype = ctypes.c_float
        _lib.blp_oracle_aten_sum.restype = ctypes.c_float
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_f32p)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(_i64p)


def num_threads():
    return lib().blp_oracle_num_threads()


def set_num_threads(n):
    lib().blp_oracle_set_num_threads(int(n))


def aten_sum(p):
    p, pp = _f(p)
    return np.float32(lib().blp_oracle_aten_sum(pp, ctypes.c_int(p.shape[-1])))


def _bcast_view(x, A, C, D):
    """(A|1, C|1, D) float32 array -> (array, strideA, strideC) in elements."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    assert x.ndim == 3 and x.shape[2] == D and x.shape[0] in (1, A) and x.shape[1] in (1, C), x.shape
    sA = 0 if x.shape[0] == 1 and A != 1 else x.shape[1] * D
    sC = 0 if x.shape[1] == 1 and C != 1 else D
    return x, sA, sC


def score_bcast(model, heads, tails, rels):
    """score_fn over 3-D operands broadcast on the two leading dims -> (A, C)."""
    D = heads.shape[-1]
    A = max(heads.shape[0], tails.shape[0], rels.shape[0])
    C = max(heads.shape[1], tails.shape[1], rels.shape[1])
    h, hA, hC = _bcast_view(heads, A, C, D)
    t, tA, tC = _bcast_view(tails, A, C, D)
    r, rA, rC = _bcast_view(rels, A, C, D)
    out = np.empty((A, C), np.float32)
    rc = lib().blp_oracle_score_bcast(
        MODELS[model], h.ctypes.data_as(_f32p), ctypes.c_int64(hA), ctypes.c_int64(hC),
        t.ctypes.data_as(_f32p), ctypes.c_int64(tA), ctypes.c_int64(tC),
        r.ctypes.data_as(_f32p), ctypes.c_int64(rA), ctypes.c_int64(rC),
        ctypes.c_int64(A), ctypes.c_int64(C), ctypes.c_int(D), out.ctypes.data_as(_f32p))
    if rc:
        raise ValueError(f"blp_oracle_score_bcast rc={rc}")
    return out


def rank_counts(pred, true_idx):
    pred, pp = _f(pred)
    ti, tp = _i(np.asarray(true_idx).reshape(-1))
    q, n = pred.shape
    gt = np.empty(q, np.int64)
    ge = np.empty(q, np.int64)
    lib().blp_oracle_rank_counts(pp, ctypes.c_int64(q), ctypes.c_int64(n), tp,
                                 gt.ctypes.data_as(_i64p), ge.ctypes.data_as(_i64p))
    return gt, ge


def metrics_from_counts(gt, ge, k_values):
    gt, gp = _i(gt)
    ge, ep = _i(ge)
    kv, kp = _i(np.asarray(k_values).reshape(-1))
    q, nk = gt.shape[0], kv.shape[0]
    recip = np.empty((q, 1), np.float32)
    hits = np.empty((q, nk), np.uint8)
    lib().blp_oracle_metrics_from_counts(gp, ep, ctypes.c_int64(q), kp, ctypes.c_int(nk),
                                         recip.ctypes.data_as(_f32p), hits.ctypes.data_as(_u8p))
    return recip, hits.astype(bool)


def eval_rank(model, ent, h_rows, t_rows, r_rows, head_idx, tail_idx,
              filt_indptr=None, filt_idx=None, want_scores=False):
    """train.py:141-171 for one batch: returns dict(gt, ge, gt_f, ge_f, true_score[, scores])."""
    ent, ep = _f(ent)
    h, hp = _f(h_rows)
    t, tp = _f(t_rows)
    r, rp = _f(r_rows)
    hi, hip = _i(head_idx) if head_idx is not None else (None, None)
    ti, tip = _i(tail_idx) if tail_idx is not None else (None, None)
    n, d = ent.shape
    b = h.shape[0]
    gt = np.empty(2 * b, np.int64); ge = np.empty(2 * b, np.int64)
    gtf = np.empty(2 * b, np.int64); gef = np.empty(2 * b, np.int64)
    ts = np.empty(2 * b, np.float32)
    scores = np.empty((2 * b, n), np.float32) if want_scores else None
    if filt_indptr is not None:
        fp_, fpp = _i(filt_indptr)
        fi_, fip = _i(filt_idx if len(filt_idx) else np.zeros(1, np.int64))
    else:
        fpp = fip = None
    rc = lib().blp_oracle_eval_rank(
        MODELS[model], ep, ctypes.c_int64(n), ctypes.c_int(d), hp, tp, rp, hip, tip, ctypes.c_int64(b),
        fpp, fip, gt.ctypes.data_as(_i64p), ge.ctypes.data_as(_i64p),
        gtf.ctypes.data_as(_i64p), gef.ctypes.data_as(_i64p), ts.ctypes.data_as(_f32p),
        scores.ctypes.data_as(_f32p) if want_scores else None)
    if rc:
        raise ValueError(f"blp_oracle_eval_rank rc={rc}")
    out = dict(gt=gt, ge=ge, gt_f=gtf, ge_f=gef, true_score=ts)
    if want_scores:
        out["scores"] = scores
    return out


def train_loss(model, loss, ent_embs, rel_rows, neg_idx, regularizer=0.0, grad_out=1.0, want_grad=True):
    """models.py:51-70 forward (+ analytic backward).  neg_idx may be any strided int64 (B,K,2) array."""
    e, ep = _f(ent_embs)
    r, rp = _f(rel_rows)
    neg_idx = np.asarray(neg_idx)
    assert neg_idx.dtype == np.int64 and neg_idx.ndim == 3 and neg_idx.shape[2] == 2
    b, k = neg_idx.shape[:2]
    d = e.shape[-1]
    s0, s1, s2 = (s // 8 for s in neg_idx.strides)
    loss_out = np.zeros(1, np.float32)
    pos = np.empty(b, np.float32)
    neg = np.empty((b, k), np.float32)
    ge_ = np.empty((b, 2, d), np.float32) if want_grad else None
    gr_ = np.empty((b, d), np.float32) if want_grad else None
    rc = lib().blp_oracle_train_loss(
        MODELS[model], LOSSES[loss], ep, rp,
        ctypes.cast(neg_idx.ctypes.data, _i64p), ctypes.c_int64(s0), ctypes.c_int64(s1), ctypes.c_int64(s2),
        ctypes.c_int64(b), ctypes.c_int64(k), ctypes.c_int(d), ctypes.c_float(regularizer),
        ctypes.c_float(grad_out), loss_out.ctypes.data_as(_f32p), pos.ctypes.data_as(_f32p),
        neg.ctypes.data_as(_f32p),
        ge_.ctypes.data_as(_f32p) if want_grad else None, gr_.ctypes.data_as(_f32p) if want_grad else None)
    if rc:
        raise ValueError(f"blp_oracle_train_loss rc={rc}")
    return dict(loss=loss_out[0], pos_scores=pos, neg_scores=neg, grad_ent=ge_, grad_rel=gr_)
