"""Tensor-level wrappers over the C ABI (include/blp_b200.h).

PyTorch is plumbing here: it owns the device memory and the stream; every
computation is a call into libblp_b200.so.  Nothing in this module computes a
score, a loss or a rank with torch ops, and nothing falls back to the CPU.
"""
import contextlib
import ctypes
import threading
import weakref

import torch

from . import _lib
from ._lib import LOSSES, MODELS, check, lib


def _require_cuda(*tensors):
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.BlpError(
                "blp_b200 runs on sm_100 CUDA tensors only (got a %s tensor); there is no CPU fallback" % t.device)
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise ValueError(f"tensors on different devices: {dev} and {t.device}")
    return dev


_checked_devices = set()


def _enter(dev):
    """Device guard + one-time architecture check; returns the current stream handle."""
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx not in _checked_devices:
        check(lib().blp_device_check(idx), "blp_device_check")
        _checked_devices.add(idx)
    return idx, ctypes.c_void_p(torch.cuda.current_stream(idx).cuda_stream)


_NULL_CTX = contextlib.nullcontext()


def _guard(dev):
    """Device guard that costs nothing when `dev` is already the current device (the common case)."""
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    return _NULL_CTX if torch.cuda.current_device() == idx else torch.cuda.device(idx)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _f32c(t):
    if t.dtype != torch.float32:
        raise ValueError(f"blp_b200 kernels are fp32 only (got {t.dtype})")
    return t.contiguous()


def model_id(name):
    try:
        return MODELS[name]
    except KeyError:
        raise ValueError(f"Unknown relational model {name}.") from None      # models.py:26


def loss_id(name):
    try:
        return LOSSES[name]
    except KeyError:
        raise ValueError(f"Unkown loss function {name}") from None           # models.py:36 (sic)


# ---------------------------------------------------------------- score_fn ----
def _as_acd(x, lead, d):
    """View operand x (broadcastable to lead + (d,)) as rows addressed by (a, c) element strides."""
    x = _f32c(x)
    if x.shape[-1] != d:
        raise ValueError(f"last dim mismatch: {tuple(x.shape)} vs d={d}")
    xl = (1,) * (len(lead) - (x.dim() - 1)) + tuple(x.shape[:-1])
    x = x.reshape(xl + (d,))
    return x, xl


def score(model, heads, tails, rels):
    """score_fn(heads, tails, rels) (models.py:222-248): broadcast leading dims, reduce the last.

    Forward only (the differentiable training path is `compute_loss`).  Operands must broadcast
    to at most two distinct leading axes, which covers every call site of the reference
    (train.py:146-147, models.py:57,67).
    """
    mid = model_id(model)
    if any(t.requires_grad for t in (heads, tails, rels)) and torch.is_grad_enabled():
        raise _lib.BlpError("blp_b200 score functions are forward-only; train through LinkPrediction.compute_loss")
    dev = _require_cuda(heads, tails, rels)
    d = heads.shape[-1]
    lead = torch.broadcast_shapes(heads.shape[:-1], tails.shape[:-1], rels.shape[:-1])
    ops, shapes = zip(*[_as_acd(x, lead, d) for x in (heads, tails, rels)])
    # collapse the broadcast lead shape to (A, C): axes [0, split) -> A, [split, nd) -> C, such that
    # every operand is either full or broadcast (all ones) on each of the two axis groups
    nd = len(lead)

    def _prod(seq):
        out = 1
        for v in seq:
            out *= v
        return out

    def _fits(sh, s):
        return all(tuple(seg) == tuple(ref) or all(v == 1 for v in seg)
                   for seg, ref in ((sh[:s], lead[:s]), (sh[s:], lead[s:])))

    split = next((s for s in range(nd + 1) if all(_fits(sh, s) for sh in shapes)), None)
    if split is None:
        raise NotImplementedError(f"broadcast pattern not supported by blp_score_bcast: {shapes}")
    A, C = _prod(lead[:split]), _prod(lead[split:])
    out = torch.empty((A, C), dtype=torch.float32, device=dev)
    args = []
    for x, sh in zip(ops, shapes):
        a_full = tuple(sh[:split]) == tuple(lead[:split])
        c_full = tuple(sh[split:]) == tuple(lead[split:])
        sA = _prod(sh[split:]) * d if (a_full and A > 1) else 0
        sC = d if (c_full and C > 1) else 0
        args += [_ptr(x), sA, sC]
    with _guard(dev):
        _, stream = _enter(dev)
        if A * C > 0:
            check(lib().blp_score_bcast(mid, *args, A, C, d, _ptr(out), stream), "blp_score_bcast")
    return out.reshape(lead)


# ------------------------------------------------------------- get_metrics ----
def rank_counts(pred_scores, true_idx):
    """Integer part of utils.get_metrics (utils.py:103-105) -> (gt, ge) int32 (Q,)."""
    dev = _require_cuda(pred_scores, true_idx)
    if pred_scores.dim() != 2:
        raise ValueError("pred_scores must be (Q, N)")
    pred = pred_scores if (pred_scores.dtype == torch.float32 and pred_scores.stride(1) == 1) else _f32c(pred_scores)
    q, n = pred.shape
    ti = true_idx.reshape(-1).to(torch.int64).contiguous()
    if ti.numel() != q:
        raise ValueError("true_idx must have one entry per row of pred_scores")
    gt = torch.empty(q, dtype=torch.int32, device=dev)
    ge = torch.empty(q, dtype=torch.int32, device=dev)
    with _guard(dev):
        _, stream = _enter(dev)
        check(lib().blp_rank_counts(_ptr(pred), q, n, pred.stride(0) if q > 1 else n, _ptr(ti), _ptr(gt), _ptr(ge), stream),
              "blp_rank_counts")
    return gt, ge


def metrics_from_counts(gt, ge, k_values):
    """Float part of utils.get_metrics (utils.py:106-109) -> (reciprocals (Q,1) f32, hits (Q,k) bool)."""
    dev = _require_cuda(gt, ge)
    ks = [int(v) for v in (k_values.reshape(-1).tolist() if torch.is_tensor(k_values) else k_values)]
    q = gt.numel()
    recip = torch.empty((q, 1), dtype=torch.float32, device=dev)
    hits = torch.empty((q, len(ks)), dtype=torch.uint8, device=dev)
    karr = (ctypes.c_int64 * max(1, len(ks)))(*ks)
    with _guard(dev):
        _, stream = _enter(dev)
        check(lib().blp_metrics_from_counts(_ptr(gt.contiguous()), _ptr(ge.contiguous()), q, karr, len(ks),
                                            _ptr(recip), _ptr(hits), stream), "blp_metrics_from_counts")
    return recip, hits.view(torch.bool)


def metrics_reduce(gt, ge, k_values):
    """train.py:154-157 accumulators -> float64 device tensor [sum 1/rank, hits@k_0, hits@k_1, ...]."""
    dev = _require_cuda(gt, ge)
    ks = [int(v) for v in (k_values.reshape(-1).tolist() if torch.is_tensor(k_values) else k_values)]
    sums = torch.empty(1 + len(ks), dtype=torch.float64, device=dev)
    karr = (ctypes.c_int64 * max(1, len(ks)))(*ks)
    gt, ge = gt.reshape(-1).contiguous(), ge.reshape(-1).contiguous()
    if gt.dtype != torch.int32 or ge.dtype != torch.int32:
        raise ValueError("rank counters must be int32")
    with _guard(dev):
        _, stream = _enter(dev)
        check(lib().blp_metrics_reduce(_ptr(gt), _ptr(ge), gt.numel(), karr, len(ks), _ptr(sums), stream),
              "blp_metrics_reduce")
    return sums


# --------------------------------------------------------- fused eval sweep ----
def eval_rank(model, ent, h_rows, t_rows, r_rows, filt_indptr=None, filt_idx=None, ent_offset=0):
    """train.py:141-171 for one batch (or many) without the (2B, N) score matrix.

    Returns dict(gt, ge[, gt_f, ge_f], true_score): int32 counts over THIS shard of the table for
    the 2B queries (head predictions then tail predictions, the reference's cat order).
    """
    mid = model_id(model)
    dev = _require_cuda(ent, h_rows, t_rows, r_rows, filt_indptr, filt_idx)
    ent, h_rows, t_rows, r_rows = (_f32c(x) for x in (ent, h_rows, t_rows, r_rows))
    if ent.dim() != 2:
        raise ValueError("ent must be (N, D)")
    n, d = ent.shape
    b = h_rows.shape[0]
    for x in (h_rows, t_rows, r_rows):
        if x.shape != (b, d):
            raise ValueError(f"query rows must be ({b}, {d}); got {tuple(x.shape)}")
    gt = torch.empty(2 * b, dtype=torch.int32, device=dev)
    ge = torch.empty(2 * b, dtype=torch.int32, device=dev)
    ts = torch.empty(2 * b, dtype=torch.float32, device=dev)
    out = {"gt": gt, "ge": ge, "true_score": ts}
    gtf = gef = None
    if filt_indptr is not None:
        filt_indptr = filt_indptr.to(torch.int64).contiguous()
        filt_idx = filt_idx.to(torch.int64).contiguous()
        if filt_indptr.numel() != 2 * b + 1:
            raise ValueError("filt_indptr must have 2B+1 entries")
        gtf = torch.empty(2 * b, dtype=torch.int32, device=dev)
        gef = torch.empty(2 * b, dtype=torch.int32, device=dev)
        out["gt_f"], out["ge_f"] = gtf, gef
    with _guard(dev):
        _, stream = _enter(dev)
        check(lib().blp_eval_rank(mid, _ptr(ent), n, int(ent_offset), d, _ptr(h_rows), _ptr(t_rows), _ptr(r_rows), b,
                                  _ptr(filt_indptr), _ptr(filt_idx), _ptr(gt), _ptr(ge), _ptr(gtf), _ptr(gef),
                                  _ptr(ts), stream), "blp_eval_rank")
    out["launches"] = _lib.last_launch_count()
    return out


_KV_LAST = threading.local()


def _kvalues(k_values):
    """Hit positions as a host list + ctypes array.  The reference passes ONE device tensor to every get_metrics call of
    an evaluation (train.py:72, 153, 167); reading it back is a device synchronisation, so the last tensor OBJECT seen by
    this thread is remembered (weak reference + version counter: an in-place change or another tensor reads again)."""
    if torch.is_tensor(k_values):
        last = getattr(_KV_LAST, "v", None)
        if last is not None and last[0]() is k_values and last[1] == k_values._version:
            ks = last[2]
        else:
            ks = [int(v) for v in k_values.reshape(-1).tolist()]
            _KV_LAST.v = (weakref.ref(k_values), k_values._version, ks)
    else:
        ks = [int(v) for v in k_values]
    return ks, (ctypes.c_int64 * max(1, len(ks)))(*ks)


def fast_table(ent):
    """Split operand table of the tensor-core mode (blp_fast_prepare_table): build once per entity table."""
    dev = _require_cuda(ent)
    if ent.dtype != torch.float32 or not ent.is_contiguous() or ent.dim() != 2:
        raise ValueError("ent must be a contiguous fp32 (N, D) tensor")
    n, d = ent.shape
    ws = torch.empty(int(lib().blp_fast_table_bytes(n)), dtype=torch.uint8, device=dev)
    with _guard(dev):
        _, stream = _enter(dev)
        check(lib().blp_fast_prepare_table(_ptr(ent), n, d, _ptr(ws), stream), "blp_fast_prepare_table")
    return ws


def refine_workspace(dev, capacity):
    """Worklist of the filter + refine mode (blp_rank_sweep_fast_exact): zeroed header {count, overflow} + `capacity`
    (query, candidate) pairs.  view(int32)[0] = entries of the last call, [1] = sticky overflow flag."""
    capacity = int(capacity)
    ws = torch.empty(int(lib().blp_fast_refine_bytes(capacity)), dtype=torch.uint8, device=dev)
    ws[:16].zero_()
    ws.refine_capacity = capacity
    return ws


def true_scores(model, ent, rel_weight, triples, out, h_rows=None, t_rows=None, ent_offset=0):
    """blp_true_scores: true-triple scores + counter reset for ALL triples of a chunked sweep, one launch."""
    mid = model_id(model)
    dev = _require_cuda(ent, rel_weight, triples, h_rows, t_rows)
    rel_weight = _f32c(rel_weight)
    n, d = ent.shape
    T = triples.shape[0]
    if h_rows is not None:
        h_rows, t_rows = _f32c(h_rows), _f32c(t_rows)
    with _guard(dev):
        _, stream = _enter(dev)
        if T > 0:
            check(lib().blp_true_scores(mid, _ptr(ent), n, int(ent_offset), d, _ptr(rel_weight), rel_weight.shape[0],
                                        _ptr(triples), T, _ptr(h_rows), _ptr(t_rows), T, _ptr(out["gt"]), _ptr(out["ge"]),
                                        _ptr(out["true_score"]), stream), "blp_true_scores")
    return _lib.last_launch_count() if T > 0 else 0


def rank_sweep_chunk(model, ent, rel_weight, triples, out, lo, hi, h_rows=None, t_rows=None, filt_indptr=None,
                     filt_idx=None, ent_offset=0, fast_table_ws=None, scores_out=None, counts_only=False,
                     refine_ws=None):
    """blp_rank_sweep on triples[lo:hi], written straight into the (2, T) arrays of `out`.

    triples (T, 3) int64 contiguous on the device: (head row, tail row, relation id); `out` holds
    contiguous (2, T) tensors gt, ge, true_score (and gt_f, ge_f when filters are given); row 0 = head
    predictions, row 1 = tail predictions.  The train.py:141-143 gathers run inside the kernels.
    """
    mid = model_id(model)
    dev = _require_cuda(ent, rel_weight, triples, h_rows, t_rows, filt_indptr, filt_idx)
    if ent.dtype != torch.float32 or not ent.is_contiguous() or ent.dim() != 2:
        raise ValueError("ent must be a contiguous fp32 (N, D) tensor")
    rel_weight = _f32c(rel_weight)
    n, d = ent.shape
    T = triples.shape[0]
    if triples.dtype != torch.int64 or not triples.is_contiguous() or triples.shape != (T, 3):
        raise ValueError("triples must be a contiguous int64 (T, 3) tensor")
    b = hi - lo
    if not (0 <= lo <= hi <= T):
        raise ValueError("bad chunk bounds")
    if (h_rows is None) != (t_rows is None):
        raise ValueError("h_rows and t_rows must both be given or both None")
    if h_rows is not None:
        h_rows, t_rows = _f32c(h_rows), _f32c(t_rows)
        if h_rows.shape != (b, d) or t_rows.shape != (b, d):
            raise ValueError(f"h_rows / t_rows must be ({b}, {d})")

    def at(name, esize=4):
        t = out.get(name)
        if t is None:
            return None
        if t.shape != (2, T) or not t.is_contiguous():
            raise ValueError(f"out[{name!r}] must be a contiguous (2, {T}) tensor")
        return ctypes.c_void_p(t.data_ptr() + lo * esize)

    if filt_indptr is not None and filt_indptr.numel() != 2 * b + 1:
        raise ValueError("filt_indptr must have 2B+1 entries")
    with _guard(dev):
        _, stream = _enter(dev)
        if b > 0:
            common = (mid, _ptr(ent), n, int(ent_offset), d, _ptr(rel_weight), rel_weight.shape[0],
                      ctypes.c_void_p(triples.data_ptr() + lo * 24), b, _ptr(h_rows), _ptr(t_rows),
                      _ptr(filt_indptr), _ptr(filt_idx), T, at("gt"), at("ge"),
                      at("gt_f") if filt_indptr is not None else None,
                      at("ge_f") if filt_indptr is not None else None, at("true_score"))
            if fast_table_ws is None and counts_only:
                check(lib().blp_rank_sweep_counts(*common, stream), "blp_rank_sweep_counts")
            elif fast_table_ws is None:
                check(lib().blp_rank_sweep(*common, stream), "blp_rank_sweep")
            elif refine_ws is not None:
                # tensor-core sweep + exact refine of the band around the true score: the exact mode's integer counters
                qws = torch.empty(int(lib().blp_fast_query_bytes(b)), dtype=torch.uint8, device=dev)
                check(lib().blp_rank_sweep_fast_exact(*common, _ptr(fast_table_ws), _ptr(qws), _ptr(refine_ws),
                                                      int(refine_ws.refine_capacity), stream), "blp_rank_sweep_fast_exact")
            else:
                # tensor-core mode (distmult / complex / simple, d = 128): tolerance-classified parity
                qws = torch.empty(int(lib().blp_fast_query_bytes(b)), dtype=torch.uint8, device=dev)
                if scores_out is not None and (scores_out.dtype != torch.float32 or scores_out.shape != (2 * b, n)
                                               or not scores_out.is_contiguous()):
                    raise ValueError(f"scores_out must be a contiguous fp32 ({2 * b}, {n}) tensor")
                check(lib().blp_rank_sweep_fast(*common, _ptr(fast_table_ws), _ptr(qws), _ptr(scores_out), n, stream),
                      "blp_rank_sweep_fast")
    return _lib.last_launch_count() if b > 0 else 0


_step_ws_lock = threading.Lock()
_step_workspaces = {}


def step_workspace(dev, stream_handle, out_len):
    """Zero-filled scratch of the fused step, one per (device, stream): the kernel leaves it zeroed (blp_b200.h)."""
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    key = (idx, stream_handle)
    nbytes = int(lib().blp_rank_step_workspace_bytes(int(out_len)))
    with _step_ws_lock:
        ws = _step_workspaces.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.zeros(max(nbytes, 1 << 16), dtype=torch.uint8, device=torch.device("cuda", idx))
            _step_workspaces[key] = ws
    return ws


def alloc_metrics(dev, out_len, nk):
    """Output tensors of the in-kernel metrics: recip (out_len, 1) f32, hits (out_len, nk) u8, sums (1 + nk) f64."""
    return {"recip": torch.empty((out_len, 1), dtype=torch.float32, device=dev),
            "hits": torch.empty((out_len, nk), dtype=torch.uint8, device=dev),
            "sums": torch.zeros(1 + nk, dtype=torch.float64, device=dev)}


def rank_step(model, ent, rel_weight, triples, out, h_rows=None, t_rows=None, ent_offset=0, group_triples=0,
              k_values=None, metrics=None):
    """blp_rank_step over ALL T triples: true scores + sweep (+ metrics) in one launch (d = 128).

    out: contiguous (2, T) tensors gt, ge (int32), true_score (f32); metrics: None, or the dict of
    `alloc_metrics(dev, 2 * T, nk)` to have utils.get_metrics / the train.py:154-157 sums computed by the same launch.
    group_triples: triples per table pass (0 = auto; the reference's eval_batch_size, e.g. 2 for Wikidata5M)."""
    mid = model_id(model)
    dev = _require_cuda(ent, rel_weight, triples, h_rows, t_rows)
    if ent.dtype != torch.float32 or not ent.is_contiguous() or ent.dim() != 2:
        raise ValueError("ent must be a contiguous fp32 (N, D) tensor")
    rel_weight = _f32c(rel_weight)
    n, d = ent.shape
    T = triples.shape[0]
    if triples.dtype != torch.int64 or not triples.is_contiguous() or triples.shape != (T, 3):
        raise ValueError("triples must be a contiguous int64 (T, 3) tensor")
    if (h_rows is None) != (t_rows is None):
        raise ValueError("h_rows and t_rows must both be given or both None")
    if h_rows is not None:
        h_rows, t_rows = _f32c(h_rows), _f32c(t_rows)
        if h_rows.shape != (T, d) or t_rows.shape != (T, d):
            raise ValueError(f"h_rows / t_rows must be ({T}, {d})")
    for name in ("gt", "ge", "true_score"):
        t_ = out[name]
        if t_.shape != (2, T) or not t_.is_contiguous():
            raise ValueError(f"out[{name!r}] must be a contiguous (2, {T}) tensor")
    ks, karr = _kvalues(k_values if k_values is not None else ())
    if metrics is not None and (metrics["recip"].numel() != 2 * T or metrics["hits"].numel() != 2 * T * len(ks)):
        raise ValueError("metrics tensors do not match (2T, nk)")
    with _guard(dev):
        _, stream = _enter(dev)
        if T == 0:
            return 0
        ws = step_workspace(dev, stream.value, 2 * T)
        check(lib().blp_rank_step(mid, _ptr(ent), n, int(ent_offset), d, _ptr(rel_weight), rel_weight.shape[0],
                                  _ptr(triples), T, _ptr(h_rows), _ptr(t_rows), T, int(group_triples),
                                  _ptr(out["gt"]), _ptr(out["ge"]), _ptr(out["true_score"]), karr, len(ks),
                                  _ptr(metrics["recip"]) if metrics else None, _ptr(metrics["hits"]) if metrics else None,
                                  _ptr(metrics["sums"]) if metrics else None, _ptr(ws), stream), "blp_rank_step")
    return _lib.last_launch_count()


def rank_queries(model, ent, hq, tq, k_values=None, ent_offset=0, want_metrics=True):
    """blp_rank_queries: rank unrelated head- / tail-prediction queries against the table without a score matrix.

    hq = (tails (n, D), rels (n, D), true (n,) int64) or None: candidate e scored as score_fn(e, tails[i], rels[i]);
    tq = (heads, rels, true) or None: score_fn(heads[i], e, rels[i]).  Returns dict(gt, ge, true_score[, recip, hits, sums])
    over the n_hq + n_tq queries, head queries first (the reference's torch.cat order, train.py:149)."""
    mid = model_id(model)
    flat = [x for part in (hq, tq) if part is not None for x in part]
    dev = _require_cuda(ent, *flat)
    if ent.dtype != torch.float32 or not ent.is_contiguous() or ent.dim() != 2:
        raise ValueError("ent must be a contiguous fp32 (N, D) tensor")
    n, d = ent.shape

    def prep(part):
        if part is None:
            return None, None, None, 0
        a, r, true = part
        a, r = _f32c(a).reshape(-1, d), _f32c(r).reshape(-1, d)
        true = true.reshape(-1).to(torch.int64).contiguous()
        if a.shape[0] != true.numel() or r.shape[0] != true.numel():
            raise ValueError("query rows and true indices must have the same length")
        return a, r, true, true.numel()

    ha, hr, ht, n_hq = prep(hq)
    ta, tr_, tt, n_tq = prep(tq)
    nq = n_hq + n_tq
    ks, karr = _kvalues(k_values if k_values is not None else ())
    buf = torch.empty((3, nq), dtype=torch.int32, device=dev)
    res = {"gt": buf[0], "ge": buf[1], "true_score": buf[2].view(torch.float32)}
    m = alloc_metrics(dev, nq, len(ks)) if want_metrics else None
    with _guard(dev):
        _, stream = _enter(dev)
        if nq > 0:
            ws = step_workspace(dev, stream.value, nq)
            check(lib().blp_rank_queries(mid, _ptr(ent), n, int(ent_offset), d, _ptr(ha), _ptr(hr), _ptr(ht), n_hq,
                                         _ptr(ta), _ptr(tr_), _ptr(tt), n_tq, _ptr(res["gt"]), _ptr(res["ge"]),
                                         _ptr(res["true_score"]), karr, len(ks), _ptr(m["recip"]) if m else None,
                                         _ptr(m["hits"]) if m else None, _ptr(m["sums"]) if m else None, _ptr(ws), stream),
                  "blp_rank_queries")
    if m:
        res.update(recip=m["recip"], hits=m["hits"].view(torch.bool), sums=m["sums"])
    res["launches"] = _lib.last_launch_count() if nq > 0 else 0
    return res


def rank_metrics(gt, ge, k_values, per_query=True):
    """utils.py:106-109 + train.py:154-157 in one launch -> (recip (Q,1) f32, hits (Q,k) bool, sums f64 [1+k])."""
    dev = _require_cuda(gt, ge)
    ks, karr = _kvalues(k_values)
    gt, ge = gt.reshape(-1), ge.reshape(-1)
    if gt.dtype != torch.int32 or ge.dtype != torch.int32 or not gt.is_contiguous() or not ge.is_contiguous():
        raise ValueError("rank counters must be contiguous int32")
    q = gt.numel()
    recip = torch.empty((q, 1), dtype=torch.float32, device=dev) if per_query else None
    hits = torch.empty((q, len(ks)), dtype=torch.uint8, device=dev) if per_query else None
    sums = torch.empty(1 + len(ks), dtype=torch.float64, device=dev)
    with _guard(dev):
        _, stream = _enter(dev)
        check(lib().blp_rank_metrics(_ptr(gt), _ptr(ge), q, karr, len(ks), _ptr(recip), _ptr(hits), _ptr(sums), stream),
              "blp_rank_metrics")
    return recip, (hits.view(torch.bool) if hits is not None else None), sums


# ------------------------------------------------ device-resident filter index ----
def filter_index_build(edges, ent2idx, n_rows, num_rel):
    """utils.get_triple_filters (utils.py:46-83) as a lookup structure built once per evaluation.

    edges (E, 3) int64 CUDA tensor of (head id, tail id, rel); ent2idx 1-D int64 CUDA tensor (id -> row or -1) or
    None when ids are rows.  Returns the opaque index workspace (uint8 tensor) for `filter_correct`."""
    dev = _require_cuda(edges, ent2idx)
    edges = edges.to(torch.int64).reshape(-1, 3).contiguous()
    e = edges.shape[0]
    if ent2idx is not None:
        ent2idx = ent2idx.to(torch.int64).contiguous()
    with _guard(dev):
        _, stream = _enter(dev)
        nbytes = int(lib().blp_filter_index_bytes(e))
        if nbytes < 0:
            check(-3, "blp_filter_index_bytes")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        check(lib().blp_filter_index_build(_ptr(edges) if e else None, e, _ptr(ent2idx),
                                           ent2idx.numel() if ent2idx is not None else 0, int(n_rows), int(num_rel),
                                           _ptr(ws), nbytes, stream), "blp_filter_index_build")
    return ws


def filter_correct(model, ent, rel_weight, triples, out, lo, hi, index_ws, num_edges, n_rows, h_rows=None, t_rows=None,
                   ent_offset=0):
    """Filtered counters of triples[lo:hi] from the raw ones already in `out` (train.py:159-167)."""
    mid = model_id(model)
    dev = _require_cuda(ent, rel_weight, triples, index_ws, h_rows, t_rows)
    rel_weight = _f32c(rel_weight)
    n, d = ent.shape
    T = triples.shape[0]
    b = hi - lo
    if b <= 0:
        return 0
    if h_rows is not None:
        h_rows, t_rows = _f32c(h_rows), _f32c(t_rows)

    def at(name):
        t = out[name]
        if t.shape != (2, T) or not t.is_contiguous():
            raise ValueError(f"out[{name!r}] must be a contiguous (2, {T}) tensor")
        return ctypes.c_void_p(t.data_ptr() + lo * 4)

    with _guard(dev):
        _, stream = _enter(dev)
        check(lib().blp_filter_correct(mid, _ptr(ent), n, int(ent_offset), d, _ptr(rel_weight), rel_weight.shape[0],
                                       ctypes.c_void_p(triples.data_ptr() + lo * 24), b, _ptr(h_rows), _ptr(t_rows),
                                       _ptr(index_ws), int(num_edges), int(n_rows), T, at("true_score"), at("gt"), at("ge"),
                                       at("gt_f"), at("ge_f"), stream), "blp_filter_correct")
    return _lib.last_launch_count()


def filter_correct_rows(model, ent, rows, h_rows, t_rows, r_rows, index, raw, gt_f, ge_f, ent_offset=0):
    """blp_filter_correct_rows: filtered counters of the 2T queries of a (head predictions | tail predictions) score
    handle from the device filter index; rows (T, 3) = (head row, tail row, relation id) are the lookup keys, the
    dense (T, D) operand rows are the call site's own gathers (train.py:141-143)."""
    mid = model_id(model)
    dev = _require_cuda(ent, rows, h_rows, t_rows, r_rows, gt_f, ge_f)
    n, d = ent.shape
    T = rows.shape[0]
    if rows.dtype != torch.int64 or not rows.is_contiguous() or rows.shape != (T, 3):
        raise ValueError("rows must be a contiguous int64 (T, 3) tensor")
    h_rows, t_rows, r_rows = (_f32c(x) for x in (h_rows, t_rows, r_rows))
    with _guard(dev):
        _, stream = _enter(dev)
        if T > 0:
            check(lib().blp_filter_correct_rows(mid, _ptr(ent), n, int(ent_offset), d, _ptr(rows), T, _ptr(h_rows), _ptr(t_rows),
                                                _ptr(r_rows), _ptr(index.workspace), index.num_edges, index.num_rows,
                                                index.num_relations, T, _ptr(raw["true_score"]), _ptr(raw["gt"]),
                                                _ptr(raw["ge"]), _ptr(gt_f), _ptr(ge_f), stream), "blp_filter_correct_rows")
    return gt_f, ge_f


def filter_correct_mask(model, ent, hq, tq, mask, raw, gt_f, ge_f, ent_offset=0):
    """blp_filter_correct_mask: filtered counters from the reference's dense (Q, N) bool mask (train.py:160-167) for the
    queries of `rank_queries(model, ent, hq, tq)`; raw = its gt / ge / true_score."""
    mid = model_id(model)
    flat = [x for part in (hq, tq) if part is not None for x in part]
    dev = _require_cuda(ent, mask, gt_f, ge_f, *flat)
    n, d = ent.shape

    def prep(part):
        if part is None:
            return None, None, None, 0
        a, r, true = part
        true = true.reshape(-1).to(torch.int64).contiguous()
        return _f32c(a).reshape(-1, d), _f32c(r).reshape(-1, d), true, true.numel()

    ha, hr, ht, n_hq = prep(hq)
    ta, tr_, tt, n_tq = prep(tq)
    if mask.dtype == torch.bool:
        mask = mask.view(torch.uint8)
    if mask.dim() != 2 or mask.shape[0] != n_hq + n_tq or mask.shape[1] < n or mask.stride(1) != 1:
        raise ValueError("mask must be a (Q, N) bool tensor with unit column stride")
    with _guard(dev):
        _, stream = _enter(dev)
        if n_hq + n_tq > 0:
            check(lib().blp_filter_correct_mask(mid, _ptr(ent), n, int(ent_offset), d, _ptr(ha), _ptr(hr), _ptr(ht), n_hq,
                                                _ptr(ta), _ptr(tr_), _ptr(tt), n_tq, _ptr(mask),
                                                mask.stride(0) if mask.shape[0] > 1 else mask.shape[1],
                                                _ptr(raw["true_score"]), _ptr(raw["gt"]), _ptr(raw["ge"]), _ptr(gt_f),
                                                _ptr(ge_f), stream), "blp_filter_correct_mask")
    return gt_f, ge_f


def mrr_breakdown(recip, triples_ids, is_new=None, rel_categories=None):
    """train.py:173-188 (utils.split_by_new_position / split_by_category) -> float64 device tensor [18]:
    mrr_by_position[3], position counts[3], mrr_by_category[2*4], category counts[4]."""
    dev = _require_cuda(recip, triples_ids, is_new, rel_categories)
    recip = _f32c(recip).reshape(-1)
    triples_ids = triples_ids.to(torch.int64).reshape(-1, 3).contiguous()
    t = triples_ids.shape[0]
    if recip.numel() != 2 * t:
        raise ValueError("recip must hold 2T reciprocal ranks (head queries, then tail queries)")
    if is_new is not None:
        is_new = is_new.to(torch.uint8).contiguous()
    if rel_categories is not None:
        rel_categories = rel_categories.to(torch.int64).contiguous()
    out = torch.empty(18, dtype=torch.float64, device=dev)
    with _guard(dev):
        _, stream = _enter(dev)
        check(lib().blp_mrr_breakdown(_ptr(recip), t, t, _ptr(triples_ids), _ptr(is_new),
                                      is_new.numel() if is_new is not None else 0, _ptr(rel_categories),
                                      rel_categories.numel() if rel_categories is not None else 0, _ptr(out), stream),
              "blp_mrr_breakdown")
    return out


# ------------------------------------------------------- negative sampler ----
def negative_sample(batch_size, num_negatives, repeats=1, *, device, seed=0, offset=0):
    """data.get_negative_sampling_indices (data.py:35-81) on the device: int64 (batch*repeats, num_negatives, 2)
    with the reference's strides (a transposed view of a (K, B*repeats, 2) buffer)."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _lib.BlpError("blp_b200 runs on sm_100 CUDA devices only; there is no CPU fallback")
    storage = torch.empty((int(num_negatives), int(batch_size) * int(repeats), 2), dtype=torch.int64, device=dev)
    with _guard(dev):
        _, stream = _enter(dev)
        check(lib().blp_negative_sample(int(batch_size), int(num_negatives), int(repeats), int(seed) & (2 ** 64 - 1),
                                        int(offset) & (2 ** 64 - 1), _ptr(storage), stream), "blp_negative_sample")
    return storage.transpose(0, 1)


# ------------------------------------------------- entity-table production ----
def store_rows(ent_shard, emb, rows=None, row0=0, normalize=False, ent_offset=0):
    """`ent_emb[idx:idx + bs] = F.normalize(batch_emb)` (models.py:38-43, train.py:95-123) into a row shard.

    ent_shard (N_local, D) fp32 contiguous: this rank's rows [ent_offset, ent_offset + N_local) of the table;
    emb (m, D): raw encoder outputs; rows: optional int64 (m,) global destination rows (default row0 + i).
    Rows owned by other ranks are skipped.  Returns ent_shard."""
    dev = _require_cuda(ent_shard, emb, rows)
    if ent_shard.dtype != torch.float32 or not ent_shard.is_contiguous() or ent_shard.dim() != 2:
        raise ValueError("ent_shard must be a contiguous fp32 (N_local, D) tensor")
    emb = _f32c(emb.detach()).reshape(-1, ent_shard.shape[1])
    if rows is not None:
        rows = rows.to(torch.int64).reshape(-1).contiguous()
        if rows.numel() != emb.shape[0]:
            raise ValueError("rows must have one entry per row of emb")
    with _guard(dev):
        _, stream = _enter(dev)
        check(lib().blp_store_rows(_ptr(emb), emb.shape[0], emb.shape[1], int(bool(normalize)), _ptr(rows), int(row0),
                                   _ptr(ent_shard), ent_shard.shape[0], int(ent_offset), stream), "blp_store_rows")
    return ent_shard


# ------------------------------------------------------ fused compute_loss ----
_ws_lock = threading.Lock()
_workspaces = {}


def _workspace(dev_idx, stream_handle, nbytes):
    """Zero-filled scratch, one per (device, stream): the kernel leaves it zeroed (blp_b200.h)."""
    key = (dev_idx, stream_handle)
    with _ws_lock:
        ws = _workspaces.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.zeros(max(nbytes, 4096), dtype=torch.uint8, device=torch.device("cuda", dev_idx))
            _workspaces[key] = ws
    return ws


def train_loss(model, loss, ent_embs, rel_weight, rels, neg_idx, regularizer=0.0, want_grad=True,
               want_neg_scores=False):
    """models.py:51-70 forward (+ analytic backward for an upstream gradient of 1) in one launch."""
    mid, lid = model_id(model), loss_id(loss)
    dev = _require_cuda(ent_embs, rel_weight, rels, neg_idx)
    ent_embs, rel_weight = _f32c(ent_embs), _f32c(rel_weight)
    if ent_embs.dim() != 3 or ent_embs.shape[1] != 2:
        raise ValueError("ent_embs must be (B, 2, D)")
    b, _, d = ent_embs.shape
    rels = rels.reshape(-1).to(torch.int64).contiguous()
    if rels.numel() != b:
        raise ValueError("rels must have B entries")
    if neg_idx.dtype != torch.int64 or neg_idx.dim() != 3 or neg_idx.shape[0] != b or neg_idx.shape[2] != 2:
        raise ValueError("neg_idx must be an int64 (B, K, 2) tensor")
    k = neg_idx.shape[1]
    s0, s1, s2 = neg_idx.stride()
    loss_out = torch.empty(1, dtype=torch.float32, device=dev)
    pos = torch.empty(b, dtype=torch.float32, device=dev)
    neg = torch.empty((b, k), dtype=torch.float32, device=dev) if want_neg_scores else None
    g_all = g_ent = g_rel = None
    if want_grad:
        # one allocation for both gradients: one memset in the library, one blp_scale in backward
        g_all = torch.empty(2 * b * d + rel_weight.numel(), dtype=torch.float32, device=dev)
        g_ent = g_all[:2 * b * d].view(b, 2, d)
        g_rel = g_all[2 * b * d:].view_as(rel_weight)
    with _guard(dev):
        idx, stream = _enter(dev)
        ws = _workspace(idx, stream.value, int(lib().blp_train_workspace_bytes(b, k)))
        check(lib().blp_train_loss(mid, lid, _ptr(ent_embs), _ptr(rel_weight), _ptr(rels), rel_weight.shape[0],
                                   _ptr(neg_idx), s0, s1, s2, b, k, d, float(regularizer), _ptr(loss_out), _ptr(pos),
                                   _ptr(neg), _ptr(g_ent), _ptr(g_rel), _ptr(ws), stream), "blp_train_loss")
    return {"loss": loss_out, "pos_scores": pos, "neg_scores": neg, "grad_ent": g_ent, "grad_rel_weight": g_rel,
            "grad_all": g_all, "launches": _lib.last_launch_count(), "workspace": ws}


def scaled(x, scale_dev):
    """x * scale_dev (device scalar) into a new tensor; the backward of the fused loss."""
    dev = _require_cuda(x, scale_dev)
    x = _f32c(x)
    scale_dev = _f32c(scale_dev.reshape(1))
    y = torch.empty_like(x)
    with _guard(dev):
        _, stream = _enter(dev)
        check(lib().blp_scale(_ptr(y), _ptr(x), _ptr(scale_dev), x.numel(), stream), "blp_scale")
    return y


def pair_loss(loss, pos_scores, neg_scores, want_grad=False):
    """loss_fn(pos_scores (B,1), neg_scores (B,K)) (models.py:251-258) -> dict(loss[, grad_pos, grad_neg])."""
    lid = loss_id(loss)
    dev = _require_cuda(pos_scores, neg_scores)
    if neg_scores.dim() != 2:
        raise ValueError("neg_scores must be (B, K)")
    b, k = neg_scores.shape
    pos = _f32c(pos_scores).reshape(-1)
    if pos.numel() != b:
        raise ValueError("pos_scores must be (B, 1)")
    neg = neg_scores if (neg_scores.dtype == torch.float32 and neg_scores.stride(1) == 1 and neg_scores.stride(0) >= k) \
        else _f32c(neg_scores)
    out = torch.empty(1, dtype=torch.float32, device=dev)
    gp = torch.empty(b, dtype=torch.float32, device=dev) if want_grad else None
    gn = torch.empty((b, k), dtype=torch.float32, device=dev) if want_grad else None
    with _guard(dev):
        _, stream = _enter(dev)
        check(lib().blp_pair_loss(lid, _ptr(pos), _ptr(neg), neg.stride(0) if b > 1 else k, b, k, _ptr(out), _ptr(gp),
                                  _ptr(gn), stream), "blp_pair_loss")
    return {"loss": out, "grad_pos": gp, "grad_neg": gn}


def l2_regularization(heads, tails, rels):
    """models.py:261-266, forward only."""
    dev = _require_cuda(heads, tails, rels)
    heads, tails, rels = (_f32c(x) for x in (heads, tails, rels))
    out = torch.empty(1, dtype=torch.float32, device=dev)
    with _guard(dev):
        _, stream = _enter(dev)
        check(lib().blp_l2_regularization(_ptr(heads), heads.numel(), _ptr(tails), tails.numel(), _ptr(rels),
                                          rels.numel(), _ptr(out), stream), "blp_l2_regularization")
    return out.reshape(())


def index_error_flag(result):
    """True if the last train_loss call on that workspace saw an out-of-range rels / neg_idx entry (syncs)."""
    flag = result["workspace"][4:8].view(torch.int32)
    bad = bool(flag.item())
    if bad:
        flag.zero_()
    return bad


def pipe_probe(variant, device, n_threads=148 * 8 * 256, iters=4096):
    """Launch one FP32 pipe micro-benchmark; returns the lane-op count (time it with CUDA events)."""
    dev = torch.device(device)
    sink = torch.empty(n_threads, dtype=torch.float32, device=dev)
    ops = ctypes.c_double(0.0)
    with _guard(dev):
        _, stream = _enter(dev)
        check(lib().blp_pipe_probe(int(variant), _ptr(sink), n_threads, iters, ctypes.byref(ops), stream), "blp_pipe_probe")
    return ops.value, sink


def atomic_probe(device, rows=2048, iters=256):
    """fp32 reduction throughput into an L2-resident (rows, 128) table (blp_atomic_probe): returns (GB/s, table)."""
    dev = torch.device(device)
    table = torch.zeros((rows, 128), dtype=torch.float32, device=dev)
    nbytes = ctypes.c_double(0.0)
    with _guard(dev):
        _, stream = _enter(dev)
        for _ in range(2):
            check(lib().blp_atomic_probe(_ptr(table), rows, iters, ctypes.byref(nbytes), stream), "blp_atomic_probe")
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            check(lib().blp_atomic_probe(_ptr(table), rows, iters, ctypes.byref(nbytes), stream), "blp_atomic_probe")
        b.record()
        torch.cuda.synchronize(dev)
    return 5 * nbytes.value / (a.elapsed_time(b) * 1e-3) / 1e9, table
