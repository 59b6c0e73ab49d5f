"""Full-entity scoring + ranking sweep: the hot loop of eval_link_prediction (train.py:128-171).

The reference scores a batch of B test triples against every candidate entity
twice (all heads, all tails), materialises a (2B, N) score matrix, and turns it
into ranks with utils.get_metrics (raw, then again after masking filtered
candidates).  Here a whole sweep is one or a few `blp_eval_rank` launches that
read the entity table once per query group and emit two integer counters per
query; reciprocal ranks / hits@k are computed once at the end.

Multi-GPU: the candidate axis is sharded by rows; each
rank counts over its shard and ONE all-reduce of the int32 counters per sweep
merges them.  Integer sums are exact, so the result is identical for any world
size.  The query rows (true head / tail rows) are exchanged once per sweep with
a bit-preserving all-reduce (every row is owned by exactly one rank).
"""
import contextlib
import ctypes

import numpy as np
import torch

from . import ops

K_VALUES = (1, 3, 10)          # train.py:66 hit_positions


def shard_bounds(n, world_size, rank):
    """Row block owned by `rank`: [lo, hi) with blocks of ceil(n / world_size) rows."""
    per = -(-int(n) // int(world_size))
    lo = min(int(n), rank * per)
    return lo, min(int(n), lo + per)


def _dist():
    import torch.distributed as dist
    return dist


def _world(group):
    """(world size, rank) of an entity-sharded sweep.  `group=None` means NOT sharded (this process holds the whole
    table) even when torch.distributed is initialised -- e.g. data-parallel replicas each ranking their own triples;
    pass `torch.distributed.group.WORLD` (or a sub-group) to shard the candidate axis over its ranks."""
    if group is None:
        return 1, 0
    dist = _dist()
    return dist.get_world_size(group), dist.get_rank(group)


_SIDE_STREAMS = {}


def _side_streams(dev):
    """Two side streams per device for chunk overlap (created once)."""
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = (torch.cuda.Stream(device=key), torch.cuda.Stream(device=key))
    return _SIDE_STREAMS[key]


def gather_rows(ent_shard, ent_offset, idx, group=None):
    """Rows `idx` (global ids) of a row-sharded table, replicated on every rank (train.py:141-142).

    Each row lives on exactly one rank; the others contribute zero bit patterns, and the sum is taken
    on the int32 view, so the result is the owner's bits exactly (including signed zeros).
    """
    world, _ = _world(group)
    idx = idx.reshape(-1)
    if world == 1:
        return ent_shard.index_select(0, idx - ent_offset if ent_offset else idx)
    local = idx - ent_offset
    mine = (local >= 0) & (local < ent_shard.shape[0])
    rows = torch.zeros((idx.numel(), ent_shard.shape[1]), dtype=ent_shard.dtype, device=ent_shard.device)
    sel = mine.nonzero().reshape(-1)
    rows.index_copy_(0, sel, ent_shard.index_select(0, local.index_select(0, sel)))
    _dist().all_reduce(rows.view(torch.int32), group=group)
    return rows


class AlignedTriples:
    """Test triples in RELATION-ALIGNED order, prepared once per evaluation set.

    The triples are sorted by relation and every relation's run is padded to a multiple of 4 entries with duplicates of
    its last triple.  The exact sweep kernel gives 4 consecutive triples to one warp; when they share the relation, the
    first rounding of head prediction -- fl(candidate + r) for TransE (models.py:223), fl(candidate * r) for the
    bilinear models (models.py:227, 235-238, 247) -- is computed once for the four of them (bit-identical results,
    ~15 % fewer FP32 lane-ops; TransE: 314 -> 284 us per 1,024 triples on FB15k-237).  Padding
    makes that hold for EVERY warp, so one loop nest stays hot.  `rank_sweep(..., triples=AlignedTriples(t))` returns
    its outputs in the caller's original order; padding entries are computed and dropped (<= 3 per relation).

        aligned = blp_b200.AlignedTriples(rows)          # rows (T, 3) int64 on the device: (head row, tail row, rel id)
        out = blp_b200.rank_sweep("transe", ent_emb, rel_weight, aligned)
    """

    def __init__(self, triples, align=4):
        if triples.dim() != 2 or triples.shape[1] != 3 or triples.dtype != torch.int64:
            raise ValueError("triples must be an int64 (T, 3) tensor")
        dev = triples.device
        T = triples.shape[0]
        self.num_triples, self.align, self.device = T, int(align), dev
        if T == 0:
            self.padded, self.src = triples.contiguous(), torch.zeros(0, dtype=torch.int64, device=dev)
            self.dest = torch.zeros(0, dtype=torch.int64, device=dev)
            return
        rel = triples[:, 2]
        perm = torch.argsort(rel, stable=True)
        srel = rel.index_select(0, perm)
        new_run = torch.ones(T, dtype=torch.bool, device=dev)
        new_run[1:] = srel[1:] != srel[:-1]
        run_id = torch.cumsum(new_run.to(torch.int64), 0) - 1
        counts = torch.zeros(T, dtype=torch.int64, device=dev).scatter_add_(0, run_id, torch.ones(T, dtype=torch.int64, device=dev))
        padded_counts = (counts + align - 1) // align * align
        start = torch.cumsum(counts, 0) - counts
        start_pad = torch.cumsum(padded_counts, 0) - padded_counts
        ar = torch.arange(T, device=dev)
        dest_sorted = start_pad.index_select(0, run_id) + (ar - start.index_select(0, run_id))
        t_pad = int(padded_counts.sum().item())                     # one host sync per evaluation set
        marks = torch.full((t_pad,), -1, dtype=torch.int64, device=dev)
        marks[dest_sorted] = torch.arange(T, device=dev)            # sorted position of every real entry
        filled = torch.cummax(marks, 0).values                      # padding entries repeat the last real entry of their run
        self.src = perm.index_select(0, filled)                     # original triple index of every padded entry
        self.padded = triples.index_select(0, self.src).contiguous()
        self.dest = torch.empty(T, dtype=torch.int64, device=dev)
        self.dest[perm] = dest_sorted                               # padded position of original triple i

    @property
    def num_padded(self):
        return self.padded.shape[0]

    def load(self, triples):
        """Refresh the padded list from new triple VALUES with the same relation column (e.g. a host copy of the same
        evaluation set arriving over PCIe): gathers into the existing buffer, no re-sorting."""
        self.padded.copy_(triples.to(self.device, non_blocking=True).index_select(0, self.src))
        return self


def rank_sweep(rel_model, ent_emb, rel_weight, triples, *, filter_index=None, filter_triples=None,
               filter_csr=None, k_values=K_VALUES, ent_offset=0, group=None, chunk=16384, group_triples=0,
               h_rows=None, t_rows=None, count_fn=None, mode="exact", fast_table=None, sort_by_relation=True,
               overlap_chunks=True):
    """Rank every test triple against all candidate entities (train.py:128-171 for the whole sweep).

    rel_model   'transe' | 'distmult' | 'complex' | 'simple'
    ent_emb     (N_local, D) fp32: this rank's row shard of the entity table (the whole table when
                not distributed); row j is candidate `ent_offset + j`
    rel_weight  (R, D) relation table (model.rel_emb.weight)
    triples     (T, 3) int64 (head_row, tail_row, rel_id): table ROWS, i.e. after ent2idx (train.py:132-135)
    filter_index / filter_triples   the filtered setting (train.py:159-167): a utils.DeviceFilterIndex (built once
                per evaluation, lookups run on the device, `filter_triples` not needed), or a host-side
                utils.TripleFilterIndex plus the (T, 3) entity-ID triples; or filter_csr = (indptr [2T+1], idx)
                precomputed with the head-prediction queries of ALL T triples first
    ent_offset / group   entity-sharded sweeps: global row id of ent_emb[0] and the process group whose ranks hold the
                other row blocks (one all-reduce of the counters per sweep); group=None = not sharded
    group_triples   exact mode, D = 128: triples per pass over the table (0 = chosen from T; 2 / 4 / 8 / 16 / 32 forces
                it).  This is what the reference's `eval_batch_size` controls -- every batch streams the whole table
                once -- so `group_triples=2` reproduces the Wikidata5M setting (HBM-bound) in ONE launch for all T
    h_rows / t_rows   optional pre-gathered (T, D) true head / tail rows (replicated)
    mode        "exact" (default): every score carries the reference's fp32 roundings, ranks are bit-exact;
                "fast": distmult / complex / simple at D = 128 as a split-FP16 tensor-core contraction
                (blp_rank_sweep_fast) -- scores within ~1e-6 * sum|terms|, ranks may differ for candidates
                inside that band around the true score; `fast_table` = ops.fast_table(ent_emb) to reuse
                the split table across calls
                "fast_exact": the tensor-core sweep with an a-priori error band around every true score; the few
                candidates inside the band are re-scored in the reference's fp32 order (blp_rank_sweep_fast_exact), so
                gt / ge / the filtered counters and every metric are bit-identical to "exact" at close to "fast" speed.
                If more candidates fall into the band than the worklist holds (degenerate tables: identical rows),
                the sweep is transparently redone in "exact" mode (`out["refine_overflow"]` tells); checking that
                costs one device synchronisation per sweep
    sort_by_relation   exact mode, D = 128, sweeps of >= 64 M scores per direction: process the triples in relation-aligned order (AlignedTriples)
                (one stable argsort per sweep; the outputs come back in the caller's order).  Triples that share a relation let the kernel compute
                fl(candidate + r) once for several head-prediction queries (~13 % fewer FP32 lane-ops); results are
                bit-identical either way
    overlap_chunks   exact sweeps cut into >= 4 chunks launch consecutive chunks on two alternating side streams
                (forked from / joined to the current stream), so one chunk's kernel start-up overlaps the previous
                chunk's tail; results are identical
    count_fn    test seam: replaces ops.eval_rank (same signature) so the sharding / collective logic can
                be exercised without a GPU

    Returns a dict: gt, ge (and gt_f, ge_f) int32 (2, T) [0 = head prediction, 1 = tail prediction],
    recip / hits (and recip_f / hits_f) per query, `sums` float64 [sum 1/rank, hits@k...] (and sums_f),
    mrr / hits_at_k (and mrr_f / hits_at_k_f) python floats normalised by 2T (train.py:196-200).
    """
    dev = ent_emb.device
    triples_in, h_rows_in, t_rows_in = triples, h_rows, t_rows         # as passed (the fast_exact overflow fallback re-runs them)
    aligned = triples if isinstance(triples, AlignedTriples) else None
    if aligned is not None:
        triples = aligned.padded
    if triples.device != dev:
        triples = triples.to(dev)
    if triples.dim() != 2:
        triples = triples.reshape(-1, 3)
    T = triples.shape[0]
    world, _ = _world(group)
    filtered = filter_index is not None or filter_csr is not None
    names = ("gt", "ge", "gt_f", "ge_f") if filtered else ("gt", "ge")
    dev_index = filter_index if hasattr(filter_index, "workspace") else None     # utils.DeviceFilterIndex

    def chunk_csr(lo, hi):
        if not filtered or dev_index is not None:
            return None, None
        if filter_csr is not None:
            indptr, idx = _slice_csr(filter_csr, lo, hi, T)
        else:
            indptr, idx = filter_index.csr(filter_triples[lo:hi])
        indptr = torch.as_tensor(indptr, dtype=torch.int64).to(dev, non_blocking=True)
        idx = torch.as_tensor(idx if len(idx) else np.zeros(1, np.int64), dtype=torch.int64).to(dev, non_blocking=True)
        return indptr, idx

    launches = 0
    if count_fn is None:
        # native path: the train.py:141-143 gathers run inside the kernels, results land in (2, T) arrays
        if triples.dtype != torch.int64 or not triples.is_contiguous():
            triples = triples.to(torch.int64).contiguous()
        if mode not in ("exact", "fast", "fast_exact"):
            raise ValueError(f"unknown mode {mode!r}")
        refine = mode == "fast_exact"
        if refine:
            mode = "fast"
        # worth one sort + a few small gathers only when the sweep itself is milliseconds long
        if (aligned is None and sort_by_relation and mode == "exact"
                and (not filtered or dev_index is not None) and ent_emb.shape[1] == 128
                and T * ent_emb.shape[0] >= 64_000_000):
            aligned = AlignedTriples(triples)
            triples = aligned.padded
            T = triples.shape[0]
        if aligned is not None:
            if filter_csr is not None or (filtered and dev_index is None):
                raise ValueError("relation-aligned triples work with a DeviceFilterIndex (or no filters), not host CSR lists")
            if h_rows is not None:          # given for the caller's T triples: bring them into the padded order
                h_rows, t_rows = h_rows.index_select(0, aligned.src), t_rows.index_select(0, aligned.src)
        if world > 1 and h_rows is None:
            h_rows = gather_rows(ent_emb, ent_offset, triples[:, 0], group)
            t_rows = gather_rows(ent_emb, ent_offset, triples[:, 1], group)
        if mode == "fast" and fast_table is None:
            fast_table = ops.fast_table(ent_emb)
            launches += 1
        # filter + refine: room for 64 band candidates per query of a chunk (a handful are expected), and never less
        # than the one block of 128 slots every epilogue warp of the grid reserves (148 x 8 x 128 = 151,552)
        refine_ws = ops.refine_workspace(dev, max(1 << 19, 128 * min(T, chunk))) if refine and T > 0 else None
        # one allocation: the int32 counters, then the fp32 true scores
        buf = torch.empty((len(names) + 1, 2, T), dtype=torch.int32, device=dev)
        counters, true_score = buf[:len(names)], buf[len(names)].view(torch.float32)
        outs = {name: counters[i] for i, name in enumerate(names)}
        outs["true_score"] = true_score
        fused_metrics = None
        # exact mode without host-side filter lists: ONE launch for the whole sweep (true scores per query group inside
        # the sweep kernel; on a single GPU without filters the same launch also produces the metrics)
        one_step = mode == "exact" and (not filtered or dev_index is not None) and ent_emb.shape[0] > 0
        if one_step:
            if world == 1 and not filtered and aligned is None:
                fused_metrics = ops.alloc_metrics(dev, 2 * T, k_values.numel() if torch.is_tensor(k_values) else len(k_values))
            launches += ops.rank_step(rel_model, ent_emb, rel_weight.detach(), triples, outs, h_rows, t_rows, ent_offset,
                                      group_triples, k_values, fused_metrics)
            if dev_index is not None:
                # filtered ranks as a sparse correction from the device-resident index: no per-batch host work
                launches += ops.filter_correct(rel_model, ent_emb, rel_weight.detach(), triples, outs, 0, T,
                                               dev_index.workspace, dev_index.num_edges, dev_index.num_rows,
                                               h_rows, t_rows, ent_offset)
        # legacy chunked path (tensor-core mode, host-built CSR filter lists): all true scores in one launch up
        # front, then one sweep launch per chunk
        split = mode == "exact" and T > chunk and not one_step
        if split:
            launches += ops.true_scores(rel_model, ent_emb, rel_weight.detach(), triples, outs, h_rows, t_rows, ent_offset)
        # consecutive chunks are independent (disjoint output slots, read-only table), so they alternate between two
        # side streams: the launch / fold / pipeline-fill head of one chunk's kernel overlaps the tail of the previous
        # one.  Fork / join with events, capture-safe.
        n_chunks = (T + chunk - 1) // chunk if T else 0
        lanes = None
        if overlap_chunks and split and n_chunks >= 4:
            main = torch.cuda.current_stream(dev)
            lanes = _side_streams(dev)
            fork = torch.cuda.Event()
            fork.record(main)
            for s_ in lanes:
                s_.wait_event(fork)
        for ci, lo in enumerate(range(0, T, chunk) if not one_step else ()):
            hi = min(T, lo + chunk)
            ctx = torch.cuda.stream(lanes[ci & 1]) if lanes is not None else contextlib.nullcontext()
            with ctx:
                indptr, idx = chunk_csr(lo, hi)         # host CSR path: its H2D copies ride the chunk's stream
                launches += ops.rank_sweep_chunk(rel_model, ent_emb, rel_weight.detach(), triples, outs, lo, hi,
                                                 None if h_rows is None else h_rows[lo:hi],
                                                 None if t_rows is None else t_rows[lo:hi], indptr, idx, ent_offset,
                                                 fast_table_ws=fast_table if mode == "fast" else None, counts_only=split,
                                                 refine_ws=refine_ws)
                if dev_index is not None:
                    launches += ops.filter_correct(rel_model, ent_emb, rel_weight.detach(), triples, outs, lo, hi,
                                                   dev_index.workspace, dev_index.num_edges, dev_index.num_rows,
                                                   None if h_rows is None else h_rows[lo:hi],
                                                   None if t_rows is None else t_rows[lo:hi], ent_offset)
        if lanes is not None:
            for s_ in lanes:
                join = torch.cuda.Event()
                join.record(s_)
                main.wait_event(join)
        if refine_ws is not None:
            overflow = refine_ws[:8].view(torch.int32)[1:2].clone()
            if world > 1:
                _dist().all_reduce(overflow, op=_dist().ReduceOp.MAX, group=group)   # every rank takes the same branch
            if int(overflow) != 0:
                # more band candidates than the worklist holds: the counters are incomplete -- redo the sweep exactly
                out = rank_sweep(rel_model, ent_emb, rel_weight, triples_in, filter_index=filter_index,
                                 filter_triples=filter_triples, filter_csr=filter_csr, k_values=k_values,
                                 ent_offset=ent_offset, group=group, chunk=chunk, group_triples=group_triples,
                                 h_rows=h_rows_in, t_rows=t_rows_in, mode="exact", sort_by_relation=sort_by_relation,
                                 overlap_chunks=overlap_chunks)
                out["refine_overflow"] = True
                out["launches"] += launches
                return out
        if aligned is not None:
            # back to the caller's order (padding entries are dropped)
            buf = buf.index_select(2, aligned.dest)
            T = aligned.num_triples
            counters, true_score = buf[:len(names)], buf[len(names)].view(torch.float32)
    else:
        # test seam: the sharding / collective logic with a CPU stand-in for blp_eval_rank
        fused_metrics = None
        heads, tails, rels = triples[:, 0].contiguous(), triples[:, 1].contiguous(), triples[:, 2].contiguous()
        if h_rows is None:
            h_rows = gather_rows(ent_emb, ent_offset, heads, group)
        if t_rows is None:
            t_rows = gather_rows(ent_emb, ent_offset, tails, group)
        r_rows = rel_weight.detach().index_select(0, rels)                  # train.py:143 rel_emb lookup
        counters = torch.zeros((len(names), 2, T), dtype=torch.int32, device=dev)
        true_score = torch.empty((2, T), dtype=torch.float32, device=dev)
        for lo in range(0, T, chunk):
            hi = min(T, lo + chunk)
            indptr, idx = chunk_csr(lo, hi)
            res = count_fn(rel_model, ent_emb, h_rows[lo:hi], t_rows[lo:hi], r_rows[lo:hi], indptr, idx, ent_offset)
            b = hi - lo
            for i, name in enumerate(names):
                counters[i, 0, lo:hi] = res[name][:b]
                counters[i, 1, lo:hi] = res[name][b:]
            true_score[0, lo:hi] = res["true_score"][:b]
            true_score[1, lo:hi] = res["true_score"][b:]
            launches += res.get("launches", 0)

    if world > 1:
        _dist().all_reduce(counters, group=group)                           # the one collective of the sweep

    out = {name: counters[i] for i, name in enumerate(names)}
    out["true_score"] = true_score
    out["launches"] = launches
    if fused_metrics is not None:
        out["recip"], out["hits"], out["sums"] = (fused_metrics["recip"], fused_metrics["hits"].view(torch.bool),
                                                  fused_metrics["sums"])
    elif count_fn is None:
        for suffix in ("", "_f") if filtered else ("",):
            recip, hits, sums = ops.rank_metrics(out["gt" + suffix], out["ge" + suffix], k_values)
            out["recip" + suffix], out["hits" + suffix], out["sums" + suffix] = recip, hits, sums
            out["launches"] += 1
    return out


class RankSweepPlan:
    """`rank_sweep` for repeated calls on the same table with a fixed number of test triples.

    Everything that does not depend on the triples is done once: argument validation, output / scratch
    allocation, and -- exact mode -- the C-side plan object (blp_plan_create) that holds every static argument, so a
    call is ONE ctypes call with three arguments and ONE kernel launch (blp_plan_run: true scores, sweep and metrics
    fused).  `overlap_calls=True` (exact mode, one GPU, no filter index) lets batch n + 1 start on the SMs batch n has
    already left -- see blp_plan_set_overlap in include/blp_b200.h for the contract on the inputs.  That matters for the reference's small eval batches (64 triples, train.py:128), where host time and
    launch count are the cost.  With a filter index the correction and the two metric reductions are separate
    launches; the tensor-core mode keeps its fold + sweep + metrics launches.  The returned tensors are STATIC: the
    next call overwrites them.

        plan = blp_b200.RankSweepPlan("transe", ent_emb, model.rel_emb.weight, num_triples=64)
        for triples in loader: out = plan(rows)      # out["sums"], out["gt"], ... as rank_sweep
    """

    def __init__(self, rel_model, ent_emb, rel_weight, num_triples, *, mode="exact", filter_index=None,
                 k_values=K_VALUES, ent_offset=0, group=None, fast_table=None, group_triples=0, overlap_calls=False):
        lib = ops.lib()
        self.model_id = ops.model_id(rel_model)
        dev = ops._require_cuda(ent_emb, rel_weight)
        if ent_emb.dtype != torch.float32 or not ent_emb.is_contiguous() or ent_emb.dim() != 2:
            raise ValueError("ent_emb must be a contiguous fp32 (N, D) tensor")
        if mode not in ("exact", "fast", "fast_exact"):
            raise ValueError(f"unknown mode {mode!r}")
        if filter_index is not None and not hasattr(filter_index, "workspace"):
            raise ValueError("RankSweepPlan takes a utils.DeviceFilterIndex (built once per evaluation)")
        self.ent, self.rel = ent_emb, ops._f32c(rel_weight.detach())
        self.dev, self.T, self.mode, self.group = dev, int(num_triples), mode, group
        self.world, _ = _world(group)
        self.ent_offset = int(ent_offset)
        n, d = ent_emb.shape
        T = self.T
        self.filtered = filter_index is not None
        self.filter_index = filter_index
        names = ("gt", "ge", "gt_f", "ge_f") if self.filtered else ("gt", "ge")
        self._plan = None
        with ops._guard(dev):
            ops._enter(dev)
            buf = torch.empty((len(names) + 1, 2, T), dtype=torch.int32, device=dev)
            self.counters = buf[:len(names)]
            self.out = {name: buf[i] for i, name in enumerate(names)}
            self.out["true_score"] = buf[len(names)].view(torch.float32)
            ks, self._karr = ops._kvalues(k_values)
            self._nk = len(ks)
            for suffix in ("", "_f") if self.filtered else ("",):
                self.out["recip" + suffix] = torch.empty((2 * T, 1), dtype=torch.float32, device=dev)
                self.out["hits" + suffix] = torch.empty((2 * T, len(ks)), dtype=torch.uint8, device=dev)
                self.out["sums" + suffix] = torch.zeros(1 + len(ks), dtype=torch.float64, device=dev)
            self.fast_table = None
            self._qws = None
            self._refine = None
            if mode in ("fast", "fast_exact"):
                self.fast_table = fast_table if fast_table is not None else ops.fast_table(ent_emb)
                self._qws = torch.empty(int(lib.blp_fast_query_bytes(T)), dtype=torch.uint8, device=dev)
            if mode == "fast_exact":
                # worklist of the refine pass; out["refine_state"] = [entries of the last call, sticky overflow flag]:
                # a non-zero flag (check it once, after the last batch: plan.refine_overflowed()) invalidates the
                # batches since the plan was created -- re-run them with mode="exact"
                self._refine = ops.refine_workspace(dev, max(1 << 19, 128 * T))
                self.out["refine_state"] = self._refine[:8].view(torch.int32)
        p = ops._ptr
        o = self.out
        # the raw metrics come out of the sweep launch itself on a single GPU (exact mode); sharded sweeps reduce
        # the counters first, so their metrics stay a separate launch
        self._fused_metrics = mode == "exact" and self.world == 1
        if mode == "exact" and T > 0:
            # private zero-invariant workspace: plans may run on different streams
            self._ws = torch.zeros(int(lib.blp_rank_step_workspace_bytes(2 * T)), dtype=torch.uint8, device=dev)
            handle = ctypes.c_void_p()
            fm = self._fused_metrics
            ops.check(lib.blp_plan_create(ctypes.byref(handle), self.model_id, p(ent_emb), n, self.ent_offset, d, p(self.rel),
                                          self.rel.shape[0], T, T, int(group_triples), p(o["gt"]), p(o["ge"]),
                                          p(o["true_score"]), self._karr, self._nk, p(o["recip"]) if fm else None,
                                          p(o["hits"]) if fm else None, p(o["sums"]) if fm else None, p(self._ws)),
                      "blp_plan_create")
            self._plan = handle
            if overlap_calls:
                # consecutive calls overlap on the GPU (programmatic dependent launch, blp_plan_set_overlap).  CONTRACT: the
                # table, the relation table and every `triples` / h_rows / t_rows passed to a call are complete when the call
                # is enqueued (resident tensors and their slices, host copies) -- not the output of a kernel enqueued just
                # before it on the same stream
                ops.check(lib.blp_plan_set_overlap(handle, 1), "blp_plan_set_overlap")
        # blp_rank_sweep_fast(model, ent, n, off, d, rel, R, TRIPLES, t, H_ROWS, T_ROWS, indptr, idx, tail_off, gt, ge, gt_f, ge_f, ts, ...)
        self._head = (self.model_id, p(ent_emb), n, self.ent_offset, d, p(self.rel), self.rel.shape[0])
        self._tail = (None, None, T, p(o["gt"]), p(o["ge"]), None, None, p(o["true_score"]))
        if mode == "fast":
            self._tail = self._tail + (p(self.fast_table), p(self._qws), None, n)
        elif mode == "fast_exact":
            self._tail = self._tail + (p(self.fast_table), p(self._qws), p(self._refine), int(self._refine.refine_capacity))
        self._lib = lib
        for suffix in ("", "_f") if self.filtered else ("",):
            o["hits" + suffix + "_bool"] = o["hits" + suffix].view(torch.bool)
        self.launches = 0

    def refine_overflowed(self):
        """fast_exact mode: did any call since the plan was created drop band candidates (device sync)?"""
        return self._refine is not None and int(self.out["refine_state"][1]) != 0

    def __del__(self):
        plan, self._plan = getattr(self, "_plan", None), None
        if plan is not None:
            try:
                self._lib.blp_plan_destroy(plan)
            except Exception:       # interpreter shutdown
                pass

    def __call__(self, triples, h_rows=None, t_rows=None):
        T, o, lib = self.T, self.out, self._lib
        if (triples.shape != (T, 3) or triples.dtype != torch.int64 or not triples.is_contiguous()
                or triples.device != self.dev):
            raise ValueError(f"triples must be a contiguous int64 ({T}, 3) tensor on {self.dev}")
        if self.world > 1 and h_rows is None:
            h_rows = gather_rows(self.ent, self.ent_offset, triples[:, 0], self.group)
            t_rows = gather_rows(self.ent, self.ent_offset, triples[:, 1], self.group)
        if h_rows is not None:
            h_rows, t_rows = ops._f32c(h_rows), ops._f32c(t_rows)
        with ops._guard(self.dev):
            _, stream = ops._enter(self.dev)
            launches = 0
            raw_metrics_done = False
            if T > 0:
                if self._plan is not None:
                    ops.check(lib.blp_plan_run(self._plan, triples.data_ptr(), ops._ptr(h_rows), ops._ptr(t_rows), stream),
                              "blp_plan_run")
                    raw_metrics_done = self._fused_metrics
                elif self._refine is not None:
                    ops.check(lib.blp_rank_sweep_fast_exact(*self._head, triples.data_ptr(), T, ops._ptr(h_rows),
                                                            ops._ptr(t_rows), *self._tail, stream), "blp_rank_sweep_fast_exact")
                else:
                    ops.check(lib.blp_rank_sweep_fast(*self._head, triples.data_ptr(), T, ops._ptr(h_rows), ops._ptr(t_rows),
                                                      *self._tail, stream), "blp_rank_sweep_fast")
                launches = ops._lib.last_launch_count()
                if self.filtered:
                    fi = self.filter_index
                    ops.check(lib.blp_filter_correct(*self._head, triples.data_ptr(), T, ops._ptr(h_rows), ops._ptr(t_rows),
                                                     ops._ptr(fi.workspace), fi.num_edges, fi.num_rows, T,
                                                     ops._ptr(o["true_score"]), ops._ptr(o["gt"]), ops._ptr(o["ge"]),
                                                     ops._ptr(o["gt_f"]), ops._ptr(o["ge_f"]), stream), "blp_filter_correct")
                    launches += 1
            if self.world > 1:
                _dist().all_reduce(self.counters, group=self.group)
            for suffix in ("", "_f") if self.filtered else ("",):
                if suffix == "" and raw_metrics_done:
                    continue
                ops.check(lib.blp_rank_metrics(ops._ptr(o["gt" + suffix]), ops._ptr(o["ge" + suffix]), 2 * T, self._karr,
                                               self._nk, ops._ptr(o["recip" + suffix]), ops._ptr(o["hits" + suffix]),
                                               ops._ptr(o["sums" + suffix]), stream), "blp_rank_metrics")
                launches += 1
        res = dict(o)
        for suffix in ("", "_f") if self.filtered else ("",):
            res["hits" + suffix] = res.pop("hits" + suffix + "_bool")
        res["launches"] = launches
        return res


def breakdowns(out, triples_ids, new_entities=None, rel_categories=None, max_ent_id=None):
    """train.py:173-188 on the device: MRR split by the position of new entities (utils.py:114-147) and by
    relation category (utils.py:150-168), from the FILTERED reciprocal ranks of a finished sweep.

    triples_ids (T, 3) entity IDS; new_entities: iterable of ids or a bool/uint8 mask; rel_categories (R,) int64.
    Returns dict(mrr_by_position (3,), mrr_pos_counts (3,), mrr_by_category (2, 4), mrr_cat_count (1, 4)) as
    float64 device tensors (the reference accumulates the same sums batch by batch in fp32)."""
    recip = out["recip_f"] if "recip_f" in out else out["recip"]
    dev = recip.device
    triples_ids = torch.as_tensor(triples_ids).to(dev)
    is_new = None
    if new_entities is not None:
        if torch.is_tensor(new_entities) and new_entities.dtype in (torch.bool, torch.uint8):
            is_new = new_entities.to(dev)
        else:
            ids = torch.as_tensor(sorted(new_entities), dtype=torch.int64)
            size = int(max(int(max_ent_id or 0), int(triples_ids[:, :2].max().item()), int(ids.max().item()) if ids.numel() else 0)) + 1
            is_new = torch.zeros(size, dtype=torch.uint8)
            is_new[ids] = 1
            is_new = is_new.to(dev)
    cats = None if rel_categories is None else torch.as_tensor(rel_categories).to(dev)
    res = ops.mrr_breakdown(recip, triples_ids, is_new, cats)
    return {"mrr_by_position": res[0:3], "mrr_pos_counts": res[3:6], "mrr_by_category": res[6:14].view(2, 4),
            "mrr_cat_count": res[14:18].view(1, 4)}


def topk_sweep(rel_model, ent_emb, rel_weight, triples, k=10, *, ent_offset=0, group=None, h_rows=None, t_rows=None,
               chunk=128, score_fn=None):
    """The k best-scoring candidate entities of every head- and tail-prediction query (the predictions behind the
    ranks of train.py:146-153), for a table that may be row-sharded over `group`.

    Every rank scores its own rows with the exact score kernels (score_fn(ent (1, N_local, D), tails, rels) /
    score_fn(heads, ent, rels): the reference's bits), keeps its k best per query, and ONE all-gather of the packed
    (score bits, global row) pairs -- 16 * k bytes per query and rank -- lets every rank merge the per-shard lists.
    Order: score descending, ties by ascending global row (a stable sort per shard, shards concatenated in rank
    order), so the result is bit-identical for any number of shards.

    triples (T, 3) int64 (head row, tail row, rel id) as for rank_sweep; h_rows / t_rows optional pre-gathered true
    rows.  Returns dict(scores (2, T, k) f32, index (2, T, k) int64 global rows; [0] = head prediction, [1] = tail
    prediction); slots beyond the number of candidates hold -inf / -1.  `score_fn` is a test seam (CPU stand-in for the
    score kernels in the gloo tests).
    """
    dev = ent_emb.device
    triples = triples.to(dev).reshape(-1, 3)
    T, k = triples.shape[0], int(k)
    if k <= 0:
        raise ValueError("k must be positive")
    world, rank = _world(group)
    n_local, d = ent_emb.shape
    if score_fn is None:
        def score_fn(heads, tails, rels):
            return ops.score(rel_model, heads, tails, rels)
    rel_rows = rel_weight.detach().index_select(0, triples[:, 2])
    if h_rows is None:
        h_rows = gather_rows(ent_emb, ent_offset, triples[:, 0], group)
    if t_rows is None:
        t_rows = gather_rows(ent_emb, ent_offset, triples[:, 1], group)
    kk = min(k, n_local)
    vals = torch.full((2, T, k), float("-inf"), dtype=torch.float32, device=dev)
    idx = torch.full((2, T, k), -1, dtype=torch.int64, device=dev)
    ent3 = ent_emb.unsqueeze(0)
    for lo in range(0, T, max(1, int(chunk))):
        hi = min(T, lo + max(1, int(chunk)))
        if kk == 0:
            break
        r = rel_rows[lo:hi].unsqueeze(1)
        for role, pred in ((0, score_fn(ent3, t_rows[lo:hi].unsqueeze(1), r)), (1, score_fn(h_rows[lo:hi].unsqueeze(1), ent3, r))):
            sv, si = torch.sort(pred, dim=1, descending=True, stable=True)     # ties: lower row first
            vals[role, lo:hi, :kk] = sv[:, :kk]
            idx[role, lo:hi, :kk] = si[:, :kk] + ent_offset
    if world > 1:
        # one collective: (score bits, global row) packed as int64 pairs, gathered from every shard
        packed = torch.stack([vals.view(torch.int32).to(torch.int64), idx], dim=-1).contiguous()
        parts = [torch.empty_like(packed) for _ in range(world)]
        _dist().all_gather(parts, packed, group=group)
        allv = torch.cat([p[..., 0].to(torch.int32).view(torch.float32) for p in parts], dim=-1)   # rank order = row order
        alli = torch.cat([p[..., 1] for p in parts], dim=-1)
        sv, order = torch.sort(allv, dim=-1, descending=True, stable=True)
        vals, idx = sv[..., :k].contiguous(), alli.gather(-1, order)[..., :k].contiguous()
    return {"scores": vals, "index": idx}


def finalize(out, num_queries=None):
    """Host-side normalisation (train.py:196-200): one D2H read of the fp64 accumulators per sweep."""
    res = {}
    for suffix in ("", "_f"):
        if "sums" + suffix not in out:
            continue
        sums = out["sums" + suffix].tolist()
        q = num_queries or out["gt"].numel()
        res["mrr" + suffix] = sums[0] / q
        res["hits_at_k" + suffix] = [s / q for s in sums[1:]]
    return res


def _slice_csr(csr, lo, hi, T):
    """Sub-CSR for triples [lo, hi): head-prediction rows lo..hi of the first T rows, then the tail rows."""
    indptr, idx = (np.asarray(a, dtype=np.int64) for a in csr)
    parts, lens = [], []
    for base in (0, T):
        s, e = indptr[base + lo], indptr[base + hi]
        parts.append(idx[s:e])
        lens.append(np.diff(indptr[base + lo:base + hi + 1]))
    sub = np.zeros(2 * (hi - lo) + 1, np.int64)
    np.cumsum(np.concatenate(lens), out=sub[1:])
    return sub, np.concatenate(parts)
