"""Score matrices that are never materialised: the fused sweep behind the reference's UNTOUCHED eval call site.

train.py:141-171 asks for two (B, N) score matrices, concatenates them, ranks them with utils.get_metrics, overwrites
the filtered entries with `pred.min() - 1.0` and ranks again:

    heads_predictions = model.score_fn(ent_emb, tail_embs, rel_embs)       # (1,N,D), (B,1,D), (B,1,D) -> (B, N)
    tails_predictions = model.score_fn(head_embs, ent_emb, rel_embs)
    pred_ents = torch.cat((heads_predictions, tails_predictions))
    reciprocals, hits = utils.get_metrics(pred_ents, true_ents, k_values)
    filters = utils.get_triple_filters(triples, filtering_graph, num_entities, ent2idx)
    filter_mask = torch.cat((heads_filter, tails_filter)).to(device)
    pred_ents[filter_mask] = pred_ents.min() - 1.0
    reciprocals, hits = utils.get_metrics(pred_ents, true_ents, k_values)

After `blp_b200.patch(models, utils)` the score functions return a `LazyScores` for that broadcast shape (no grad):
a torch.Tensor subclass with the right shape / dtype / device and no storage, which remembers the queries.  torch.cat of
such handles is again a handle; `get_metrics` on a handle runs ONE fused launch (blp_rank_queries: true scores, sweep,
counters, reciprocal ranks, hits); `pred[mask] = pred.min() - 1.0` is recorded, not executed, and the second
get_metrics derives the filtered ranks as a sparse correction of the first (blp_filter_correct_rows from the
device-resident filter index when utils.get_triple_filters is patched too, blp_filter_correct_mask from the
reference's dense mask otherwise).  Every other operation on a handle materialises the matrix with the exact
score kernels and proceeds on the real tensor, so nothing a caller could do with the reference's return value breaks.
"""
import weakref

import torch

from . import ops

_ENABLED = {"scores": False}


def enable(flag=True):
    _ENABLED["scores"] = bool(flag)


def enabled():
    return _ENABLED["scores"]


def _meta_funcs():
    t = torch.Tensor
    fs = {t.size, t.dim, t.numel, t.__len__, t.is_contiguous, t.stride, t.storage_offset, t.is_floating_point,
          t.is_complex, t.get_device, t.element_size, t.nelement, t.ndimension, t.type}
    for name in ("shape", "device", "dtype", "ndim", "requires_grad", "is_cuda", "layout", "grad", "grad_fn", "is_leaf",
                 "is_sparse", "is_quantized", "is_meta", "names", "_version"):
        prop = getattr(t, name, None)
        if prop is not None and hasattr(prop, "__get__"):
            fs.add(prop.__get__)
    return fs


_META = None


class LazyMin:
    """`pred.min()` of a handle, and `pred.min() - c`: only meaningful as the fill value of the reference's filter
    statement (train.py:165); anything else (float(), .item(), arithmetic with tensors) computes the real minimum."""

    def __init__(self, src, offset=0.0):
        self.src, self.offset = src, float(offset)

    def __sub__(self, c):
        if isinstance(c, (int, float)):
            return LazyMin(self.src, self.offset - float(c))
        return self.value() - c

    def __add__(self, c):
        if isinstance(c, (int, float)):
            return LazyMin(self.src, self.offset + float(c))
        return self.value() + c

    __radd__ = __add__

    def value(self):
        return self.src.materialize().min() + self.offset

    def item(self):
        return self.value().item()

    def __float__(self):
        return float(self.value())


class _Part:
    """`n` queries of one role: head prediction scores candidate e as score_fn(e, rows_a, rows_r) (train.py:146),
    tail prediction as score_fn(rows_a, e, rows_r) (train.py:147)."""
    __slots__ = ("head_pred", "rows_a", "rows_r", "n")

    def __init__(self, head_pred, rows_a, rows_r):
        self.head_pred, self.rows_a, self.rows_r, self.n = head_pred, rows_a, rows_r, rows_a.shape[0]


class LazyScores(torch.Tensor):
    @staticmethod
    def __new__(cls, model, ent, parts):
        q = sum(p.n for p in parts)
        r = torch.Tensor._make_wrapper_subclass(cls, (q, ent.shape[0]), dtype=torch.float32, device=ent.device,
                                                requires_grad=False)
        r._model, r._ent, r._parts = model, ent, list(parts)
        r._dense = None          # the real tensor once something needed it
        r._raw = None            # cached raw counters of the first get_metrics
        r._true_ptr = None
        r._filter = None         # recorded `pred[mask] = pred.min() - c`
        return r

    # ------------------------------------------------------------------ torch plumbing ----
    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        global _META
        kwargs = kwargs or {}
        if _META is None:
            _META = _meta_funcs()
        if func in _META:
            with torch._C.DisableTorchFunctionSubclass():
                return func(*args, **kwargs)
        if func is torch.cat:
            tensors = args[0] if args else kwargs.get("tensors")
            dim = args[1] if len(args) > 1 else kwargs.get("dim", 0)
            if (dim in (0, -2) and len(tensors) > 0 and all(isinstance(x, LazyScores) and x._dense is None and x._filter is None
                                                             for x in tensors)
                    and all(_same_table(x._ent, tensors[0]._ent) and x._model == tensors[0]._model for x in tensors)):
                return LazyScores(tensors[0]._model, tensors[0]._ent, [p for x in tensors for p in x._parts])
        if func is torch.Tensor.min and len(args) == 1 and not kwargs and args[0]._dense is None:
            return LazyMin(args[0])
        if func is torch.Tensor.__setitem__ and len(args) == 3 and isinstance(args[0], LazyScores) and args[0]._dense is None:
            self, mask, value = args
            if (isinstance(value, LazyMin) and value.src is self and value.offset < 0 and self._filter is None
                    and _is_mask_for(mask, self)):
                self._filter = mask
                return None
        return func(*_materialized(args), **_materialized(kwargs))

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        return func(*_materialized(args), **_materialized(kwargs or {}))

    def __repr__(self):
        state = "materialised" if self._dense is not None else "lazy"
        return f"LazyScores({self._model}, shape={tuple(self.shape)}, {state})"

    def materialize(self):
        """The real (Q, N) score matrix (exact score kernels, bit-identical to the reference's values)."""
        if self._dense is None:
            ent3 = self._ent.unsqueeze(0)
            rows = []
            for p in self._parts:
                a, r = p.rows_a.unsqueeze(1), p.rows_r.unsqueeze(1)
                rows.append(ops.score(self._model, ent3, a, r) if p.head_pred else ops.score(self._model, a, ent3, r))
            dense = torch.cat(rows) if len(rows) != 1 else rows[0]
            if self._filter is not None:        # the recorded train.py:165 statement
                mask = self._filter.materialize() if isinstance(self._filter, LazyFilter) else self._filter
                dense[mask.to(dense.device)] = dense.min() - 1.0
            self._dense = dense
        return self._dense

    # ------------------------------------------------------------------ get_metrics ----
    def metrics(self, true_idx, k_values):
        """utils.get_metrics(self, true_idx, k_values) (utils.py:86-111) without the matrix."""
        dev = self._ent.device
        true_idx = true_idx.reshape(-1).to(dev)
        if true_idx.numel() != self.shape[0]:
            raise ValueError("true_idx must have one entry per row of pred_scores")
        key = (true_idx.data_ptr(), true_idx._version)
        if self._raw is None or self._true_ptr != key:
            self._raw = self._rank(true_idx, k_values)
            self._true_ptr = key
            self._true_keep = true_idx          # keeps the storage alive: its address cannot be handed to another tensor
        raw = self._raw
        if self._filter is None:
            return raw["recip"], raw["hits"]
        gt_f, ge_f = self._filtered_counts(raw, true_idx)
        recip, hits, _ = ops.rank_metrics(gt_f, ge_f, k_values)
        return recip, hits

    def _pattern(self):
        """(head part, tail part) when the handle is cat(head predictions, tail predictions), else None."""
        ps = self._parts
        if len(ps) == 2 and ps[0].head_pred and not ps[1].head_pred:
            return ps[0], ps[1]
        return None

    def _rank(self, true_idx, k_values):
        pat = self._pattern()
        if pat is not None:
            hp, tp = pat
            return ops.rank_queries(self._model, self._ent, (hp.rows_a, hp.rows_r, true_idx[:hp.n]),
                                    (tp.rows_a, tp.rows_r, true_idx[hp.n:]), k_values=k_values)
        # any other arrangement of parts: one single-role pass per part, metrics over the concatenation
        outs, lo = [], 0
        for p in self._parts:
            q = (p.rows_a, p.rows_r, true_idx[lo:lo + p.n])
            outs.append(ops.rank_queries(self._model, self._ent, q if p.head_pred else None, None if p.head_pred else q,
                                         want_metrics=False))
            lo += p.n
        res = {k: torch.cat([o[k] for o in outs]) for k in ("gt", "ge", "true_score")}
        res["recip"], res["hits"], res["sums"] = ops.rank_metrics(res["gt"], res["ge"], k_values)
        return res

    def _filtered_counts(self, raw, true_idx):
        dev = self._ent.device
        q = self.shape[0]
        gt_f = torch.empty(q, dtype=torch.int32, device=dev)
        ge_f = torch.empty(q, dtype=torch.int32, device=dev)
        pat = self._pattern()
        flt = self._filter
        if (isinstance(flt, LazyFilter) and pat is not None and flt.both and pat[0].n == pat[1].n == flt.num_triples
                and pat[0].rows_r.data_ptr() == pat[1].rows_r.data_ptr()):
            hp, tp = pat
            idx = flt.index(dev)
            ops.filter_correct_rows(self._model, self._ent, flt.rows(dev), tp.rows_a, hp.rows_a, hp.rows_r, idx, raw, gt_f, ge_f)
            return gt_f, ge_f
        mask = flt.materialize() if isinstance(flt, LazyFilter) else flt
        mask = mask.to(dev)
        if mask.dtype != torch.bool or tuple(mask.shape) != tuple(self.shape):
            raise ValueError("filter mask must be a bool tensor with the shape of the score matrix")
        lo = 0
        for p in self._parts:       # the mask kernel takes (head queries, tail queries); feed it part by part
            sl = slice(lo, lo + p.n)
            qd = (p.rows_a, p.rows_r, true_idx[sl])
            ops.filter_correct_mask(self._model, self._ent, qd if p.head_pred else None, None if p.head_pred else qd,
                                    mask[sl], {k: raw[k][sl] for k in ("gt", "ge", "true_score")}, gt_f[sl], ge_f[sl])
            lo += p.n
        return gt_f, ge_f


def _same_table(a, b):
    return a is b or (a.data_ptr() == b.data_ptr() and a.shape == b.shape and a.stride() == b.stride() and a.device == b.device)


def _materialized(x):
    if isinstance(x, LazyScores):
        return x.materialize()
    if isinstance(x, LazyFilter):
        return x.materialize()
    if isinstance(x, LazyMin):
        return x.value()
    if isinstance(x, (list, tuple)):
        return type(x)(_materialized(v) for v in x)
    if isinstance(x, dict):
        return {k: _materialized(v) for k, v in x.items()}
    return x


def _is_mask_for(mask, scores):
    if isinstance(mask, LazyFilter):
        return tuple(mask.shape) == tuple(scores.shape)
    return torch.is_tensor(mask) and mask.dtype == torch.bool and tuple(mask.shape) == tuple(scores.shape)


def maybe_lazy_score(model, heads, tails, rels):
    """The handle for an eval-shaped broadcast (train.py:146-147), or None when the call is anything else."""
    if not _ENABLED["scores"] or torch.is_grad_enabled() and any(t.requires_grad for t in (heads, tails, rels)):
        return None
    if not (heads.is_cuda and tails.is_cuda and rels.is_cuda) or heads.dim() != 3 or tails.dim() != 3 or rels.dim() != 3:
        return None
    if any(t.dtype != torch.float32 for t in (heads, tails, rels)):
        return None
    d = heads.shape[-1]
    for cand, other, head_pred in ((heads, tails, True), (tails, heads, False)):
        if (cand.shape[0] == 1 and cand.shape[1] > 1 and other.shape[1] == 1 and rels.shape[1] == 1
                and other.shape[0] == rels.shape[0] and other.shape[0] >= 1 and cand.is_contiguous()
                and other.shape[-1] == d and rels.shape[-1] == d and cand.shape[-1] == d):
            ent = cand[0]
            return LazyScores(model, ent, [_Part(head_pred, other.reshape(-1, d).contiguous(), rels.reshape(-1, d).contiguous())])
    return None


# ------------------------------------------------------------------------ filters ----
_INDEX_CACHE = {}


def _index_for(graph, ent2idx, num_ents, dev):
    """One DeviceFilterIndex per (filtering graph, ent2idx CONTENT, device), built on first use (utils.py:46-83 per
    batch in the reference).  The graph must not change while it is used for filtering (train.py never changes it).

    ent2idx is rebuilt by every evaluation (train.py:87) and differs between the validation and the test evaluation
    over the same graph, while a freed tensor's address is readily handed out again -- so an entry is matched by
    OBJECT (weak reference + version counter: the per-batch calls of one evaluation, no data touched) and, for an object
    not seen before, by comparing its content with the copy the entry was built from (once per evaluation)."""
    from .utils import DeviceFilterIndex, graph_edges
    key = (id(graph), int(num_ents), str(dev))
    entries = _INDEX_CACHE.get(key)
    if entries is not None and entries[0]["graph"]() is not graph:
        entries = None                                     # a recycled id(): not the graph the entries were built for
    if entries:
        for e in entries:
            if e["e2i_ref"]() is ent2idx and e["version"] == ent2idx._version:
                return e["idx"], e["e2i_dev"]
        host = ent2idx.detach().cpu()
        for e in entries:
            if e["e2i_host"].shape == host.shape and torch.equal(e["e2i_host"], host):
                e["e2i_ref"], e["version"] = weakref.ref(ent2idx), ent2idx._version
                return e["idx"], e["e2i_dev"]
    edges = graph_edges(graph)
    num_rel = int(edges[:, 2].max()) + 1 if len(edges) else 1
    idx = DeviceFilterIndex(edges, ent2idx, num_ents, num_rel, dev)
    entry = {"graph": weakref.ref(graph, lambda _r, k=key: _INDEX_CACHE.pop(k, None)), "e2i_ref": weakref.ref(ent2idx),
             "version": ent2idx._version, "e2i_host": ent2idx.detach().cpu().clone(), "idx": idx, "e2i_dev": ent2idx.to(dev)}
    if entries is None:
        _INDEX_CACHE[key] = entries = []
    entries.insert(0, entry)
    del entries[4:]                                        # validation + test (+ spare) per graph
    return idx, entry["e2i_dev"]


class LazyFilter(torch.Tensor):
    """utils.get_triple_filters' return value ((B, N) bool masks, utils.py:46-83) as a handle: the dense mask is only
    built (by the reference's own function) if something other than the train.py:164-165 statements touches it."""

    @staticmethod
    def __new__(cls, triples, graph, num_ents, ent2idx, roles, orig_fn):
        b = triples.shape[0]
        r = torch.Tensor._make_wrapper_subclass(cls, (b * len(roles), num_ents), dtype=torch.bool, device=torch.device("cpu"),
                                                requires_grad=False)
        r._triples, r._graph, r._num_ents, r._ent2idx, r._roles, r._orig = triples, graph, num_ents, ent2idx, roles, orig_fn
        r._dense = None
        return r

    @property
    def both(self):
        return self._roles == ("head", "tail")

    @property
    def num_triples(self):
        return self._triples.shape[0]

    def index(self, dev):
        return _index_for(self._graph, self._ent2idx, self._num_ents, dev)[0]

    def rows(self, dev):
        """(B, 3) int64 on the device: (head row, tail row, relation id) -- the lookup keys of the filter index."""
        tr, e2i_host = self._triples, self._ent2idx
        if not tr.is_cuda and torch.is_tensor(e2i_host) and not e2i_host.is_cuda:
            # the reference's layout (train.py:87-93, 133-135: CPU triples, CPU ent2idx): map on the host, ONE small H2D
            # copy per batch instead of a copy + two device gathers + a stack
            tr = tr.to(torch.int64)
            return torch.stack([e2i_host[tr[:, 0]], e2i_host[tr[:, 1]], tr[:, 2]], dim=1).to(dev)
        _, e2i = _index_for(self._graph, self._ent2idx, self._num_ents, dev)
        tr = tr.to(dev)
        return torch.stack([e2i[tr[:, 0]], e2i[tr[:, 1]], tr[:, 2]], dim=1).contiguous()

    def materialize(self):
        if self._dense is None:
            hf, tf = self._orig(self._triples, self._graph, self._num_ents, self._ent2idx)
            self._dense = torch.cat([hf if r == "head" else tf for r in self._roles])
        return self._dense

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        global _META
        kwargs = kwargs or {}
        if _META is None:
            _META = _meta_funcs()
        if func in _META:
            with torch._C.DisableTorchFunctionSubclass():
                return func(*args, **kwargs)
        if func is torch.cat:
            tensors = args[0] if args else kwargs.get("tensors")
            dim = args[1] if len(args) > 1 else kwargs.get("dim", 0)
            if (dim in (0, -2) and len(tensors) == 2 and all(isinstance(x, LazyFilter) and x._dense is None for x in tensors)
                    and tensors[0]._triples is tensors[1]._triples and tensors[0]._roles == ("head",)
                    and tensors[1]._roles == ("tail",)):
                a = tensors[0]
                return LazyFilter(a._triples, a._graph, a._num_ents, a._ent2idx, ("head", "tail"), a._orig)
        if func is torch.Tensor.to and isinstance(args[0], LazyFilter) and args[0]._dense is None:
            return args[0]                      # `.to(device)` (train.py:164): the index already lives on the device
        return func(*_materialized(args), **_materialized(kwargs))

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        return func(*_materialized(args), **_materialized(kwargs or {}))

    def __repr__(self):
        return f"LazyFilter(roles={self._roles}, shape={tuple(self.shape)})"


def make_get_triple_filters(orig_fn):
    """Drop-in for utils.get_triple_filters (utils.py:46-83): handles instead of dense masks; `orig_fn` (the
    reference's own function) builds the real masks if a caller needs them."""

    def get_triple_filters(triples, graph, num_ents, ent2idx):
        if not _ENABLED["scores"] or not torch.cuda.is_available():
            return orig_fn(triples, graph, num_ents, ent2idx)
        return (LazyFilter(triples, graph, num_ents, ent2idx, ("head",), orig_fn),
                LazyFilter(triples, graph, num_ents, ent2idx, ("tail",), orig_fn))

    get_triple_filters.__doc__ = orig_fn.__doc__
    get_triple_filters.__wrapped__ = orig_fn
    return get_triple_filters
