"""Host-side logic on CPU: filter CSR construction (utils.py:46-83), row sharding, the
sweep's chunking / collective plumbing (world_size 2 over gloo), and the end-to-end
eval metrics against the reference's eval_link_prediction run (golden eval_loop_*).
The per-shard counting is done by the ORACLE here (count_fn seam) -- the CUDA kernels
are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch

from conftest import golden
from oracle import c_oracle, np_oracle

import blp_b200
from blp_b200.evaluate import _slice_csr, rank_sweep, shard_bounds
from blp_b200.utils import TripleFilterIndex, make_ent2idx

MODELS = ("transe", "distmult", "complex", "simple")


def oracle_count_fn(model, ent, h_rows, t_rows, r_rows, indptr, idx, ent_offset):
    """Stand-in for ops.eval_rank on CPU tensors: counts over THIS shard, filter ids are global."""
    n_local = ent.shape[0]
    b = h_rows.shape[0]
    fi = fp = None
    if indptr is not None:
        indptr, idx = indptr.numpy(), idx.numpy()
        lists = []
        for q in range(2 * b):
            cols = idx[indptr[q]:indptr[q + 1]] - ent_offset
            lists.append(cols[(cols >= 0) & (cols < n_local)])
        fp = np.concatenate([[0], np.cumsum([len(x) for x in lists])]).astype(np.int64)
        fi = np.concatenate(lists) if fp[-1] else np.zeros(0, np.int64)
    if n_local == 0:
        z = np.zeros(2 * b, np.int64)
        out = dict(gt=z, ge=z, gt_f=z, ge_f=z, true_score=np.zeros(2 * b, np.float32))
    else:
        out = c_oracle.eval_rank(model, ent.numpy(), h_rows.numpy(), t_rows.numpy(), r_rows.numpy(), None, None, fp, fi)
    res = {k: torch.from_numpy(np.asarray(v)).to(torch.int32) for k, v in out.items() if k != "true_score"}
    res["true_score"] = torch.from_numpy(out["true_score"])
    return res


def reference_filters(triples, edges, num_ents, ent2idx):
    """utils.get_triple_filters restated literally over an edge list (utils.py:46-83)."""
    hf = np.zeros((len(triples), num_ents), bool)
    tf = np.zeros_like(hf)
    for i, (head, tail, rel) in enumerate(triples.tolist()):
        for h, t, r in edges.tolist():
            if h == head and r == rel and t != tail and ent2idx[t] != -1:
                tf[i, ent2idx[t]] = True
            if t == tail and r == rel and h != head and ent2idx[h] != -1:
                hf[i, ent2idx[h]] = True
    return hf, tf


def test_make_ent2idx_docstring_case():
    assert make_ent2idx(torch.tensor([4, 5, 0]), 5).tolist() == [2, -1, -1, -1, 0, 1]     # utils.py:36-38


def test_shard_bounds_cover_and_disjoint():
    for n in (0, 1, 7, 135, 14541, 4800000):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_filter_index_matches_reference_semantics():
    g = golden("eval_loop_transe")
    ent2idx = make_ent2idx(torch.from_numpy(g["entities"]), int(g["n_ids"]) - 1).numpy()
    fidx = TripleFilterIndex(g["graph_edges"], ent2idx)
    n = len(g["entities"])
    hf, tf = fidx.dense_masks(g["triples"], n)
    rhf, rtf = reference_filters(g["triples"], g["graph_edges"], n, ent2idx)
    assert np.array_equal(hf, rhf) and np.array_equal(tf, rtf)
    assert rhf.sum() + rtf.sum() > 0
    # chunk slicing of a whole-sweep CSR
    indptr, idx = fidx.csr(g["triples"])
    T = len(g["triples"])
    sub_ptr, sub_idx = _slice_csr((indptr, idx), 16, 48, T)
    ref_ptr, ref_idx = fidx.csr(g["triples"][16:48])
    assert np.array_equal(sub_ptr, ref_ptr) and np.array_equal(sub_idx, ref_idx)


def _loop_inputs(model):
    g = golden("eval_loop_" + model)
    ent2idx = make_ent2idx(torch.from_numpy(g["entities"]), int(g["n_ids"]) - 1)
    triples = torch.from_numpy(g["triples"])
    rows = torch.stack([ent2idx[triples[:, 0]], ent2idx[triples[:, 1]], triples[:, 2]], dim=1)
    fidx = TripleFilterIndex(g["graph_edges"], ent2idx)
    return g, rows, fidx


def _metrics(out, T):
    res = {}
    for suffix, tag in (("", ""), ("_f", "_filt")):
        gt, ge = out["gt" + suffix].reshape(-1).numpy(), out["ge" + suffix].reshape(-1).numpy()
        recip, hits = c_oracle.metrics_from_counts(gt, ge, [1, 3, 10])
        res["test_mrr" + tag] = float(recip.astype(np.float64).sum() / (2 * T))
        for j, k in enumerate((1, 3, 10)):
            res[f"test_hits@{k}{tag}"] = float(hits[:, j].sum() / (2 * T))
        res["recip" + suffix] = recip.reshape(2, T)
    return res


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("chunk", (16, 40, 1000))
def test_sweep_reproduces_reference_eval_loop(model, chunk):
    g, rows, fidx = _loop_inputs(model)
    T = rows.shape[0]
    out = rank_sweep(model, torch.from_numpy(g["ent_emb"]), torch.from_numpy(g["rel_weight"]), rows,
                     filter_index=fidx, filter_triples=g["triples"], chunk=chunk, count_fn=oracle_count_fn)
    m = _metrics(out, T)
    want = dict(zip(g["scalar_names"].tolist(), g["scalar_values"].tolist()))
    for name in ("test_mrr", "test_mrr_filt", "test_hits@1", "test_hits@3", "test_hits@10",
                 "test_hits@1_filt", "test_hits@3_filt", "test_hits@10_filt"):
        assert abs(m[name] - want[name]) <= 1e-6, (name, m[name], want[name])
    # by-position breakdown (utils.py:114-148) from the per-query reciprocals
    new = set(g["new_entities"].tolist())
    rf = m["recip_f"]
    sums, cnts = np.zeros(3), np.zeros(3)
    for i, (h, t, _) in enumerate(g["triples"].tolist()):
        v = (float(rf[0, i]) + float(rf[1, i])) / 2.0
        slot = 0 if (h in new and t in new) else 1 if h in new else 2 if t in new else None
        if slot is not None:
            sums[slot] += v
            cnts[slot] += 1
    cnts[cnts < 1] = 1
    for slot, name in enumerate(("test_mrr_filt_both_new", "test_mrr_filt_head_new", "test_mrr_filt_tail_new")):
        assert abs(sums[slot] / cnts[slot] - want[name]) <= 1e-5, name


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, model, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g, rows, fidx = _loop_inputs(model)
        table = torch.from_numpy(g["ent_emb"])
        lo, hi = shard_bounds(table.shape[0], world, rank)
        out = rank_sweep(model, table[lo:hi].contiguous(), torch.from_numpy(g["rel_weight"]), rows,
                         filter_index=fidx, filter_triples=g["triples"], chunk=40, ent_offset=lo,
                         group=dist.group.WORLD, count_fn=oracle_count_fn)
        if rank == 0:
            ret.update({k: out[k].numpy() for k in ("gt", "ge", "gt_f", "ge_f")})
        # group=None inside an initialised process group = NOT sharded (data-parallel replicas, bench.py --gpus N):
        # no collective, every rank ranks its own triples against its own full table
        mine = rows[rank::world]
        local = rank_sweep(model, table, torch.from_numpy(g["rel_weight"]), mine, count_fn=oracle_count_fn)
        ret[f"replica_gt_{rank}"] = local["gt"].numpy()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("model", ("transe", "complex"))
@pytest.mark.parametrize("world", (2, 3))
def test_sharded_sweep_equals_single_rank_gloo(model, world):
    import torch.multiprocessing as mp
    g, rows, fidx = _loop_inputs(model)
    single = rank_sweep(model, torch.from_numpy(g["ent_emb"]), torch.from_numpy(g["rel_weight"]), rows,
                        filter_index=fidx, filter_triples=g["triples"], count_fn=oracle_count_fn)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), model, ret), nprocs=world, join=True)
    for k in ("gt", "ge", "gt_f", "ge_f"):
        assert np.array_equal(ret[k], single[k].numpy()), k      # integer sums: bit-identical for any world size
    for rank in range(world):
        assert np.array_equal(ret[f"replica_gt_{rank}"], single["gt"].numpy()[:, rank::world]), rank


def test_gather_rows_single_process_matches_index():
    table = torch.randn(50, 8)
    idx = torch.tensor([3, 49, 0, 3])
    assert torch.equal(blp_b200.gather_rows(table, 0, idx), table[idx])


def test_torch_port_row_restatements_are_consistent():
    """The CPU timing baselines of the rows either side of the path (oracle/torch_port.py) agree with the host-side
    index logic they are timed against: filter masks == TripleFilterIndex.dense_masks; sampler structure (data.py:35-81)."""
    from oracle import torch_port
    rng = np.random.default_rng(3)
    n_ids, n_rel = 60, 4
    edges = np.stack([rng.integers(0, n_ids, 500), rng.integers(0, n_ids, 500), rng.integers(0, n_rel, 500)], 1)
    entities = torch.from_numpy(rng.permutation(n_ids)[:45].astype(np.int64))
    ent2idx = blp_b200.make_ent2idx(entities, n_ids - 1)
    triples = torch.from_numpy(np.stack([entities.numpy()[rng.integers(0, 45, 20)], entities.numpy()[rng.integers(0, 45, 20)],
                                         rng.integers(0, n_rel, 20)], 1))
    out_edges, in_edges = {}, {}
    for h, t, r in edges.tolist():
        out_edges.setdefault(h, []).append((h, t, r))
        in_edges.setdefault(t, []).append((h, t, r))
    hm, tm = torch_port.triple_filter_masks(triples, out_edges, in_edges, 45, ent2idx.tolist())
    want_h, want_t = blp_b200.TripleFilterIndex(edges, ent2idx).dense_masks(triples, 45)
    assert np.array_equal(hm.numpy(), want_h) and np.array_equal(tm.numpy(), want_t)
    neg = torch_port.sample_negative_indices(8, 50)
    assert tuple(neg.shape) == (8, 50, 2) and tuple(neg.stride()) == (2, 16, 1)
    own = torch.arange(16).view(8, 1, 2)
    keep = neg == own
    assert bool((keep.sum(-1) == 1).all())                                  # exactly one side keeps the own entity
    repl = torch.where(keep[..., 0], neg[..., 1], neg[..., 0])
    assert bool(((repl // 2) != torch.arange(8).view(8, 1)).all())          # the other comes from another row
    x = torch.randn(5, 128)
    assert np.array_equal(torch_port.normalize_rows(x).numpy(), np_oracle.l2_normalize_rows(x.numpy()))


def test_filter_index_from_networkx_graph_as_integration_md():
    """INTEGRATION.md Level 1 builds the index from the reference's nx.MultiDiGraph (train.py:298-302:
    add_weighted_edges_from stores the relation as the edge WEIGHT); the result must equal utils.get_triple_filters
    restated literally over that graph (utils.py:46-83), including parallel edges with different relations."""
    nx = pytest.importorskip("networkx")
    rng = np.random.default_rng(5)
    n_ids, n_rel, n_edges = 40, 4, 400
    all_triples = np.stack([rng.integers(0, n_ids, n_edges), rng.integers(0, n_ids, n_edges),
                            rng.integers(0, n_rel, n_edges)], axis=1)
    all_triples[1] = all_triples[0]; all_triples[1, 2] = (all_triples[0, 2] + 1) % n_rel    # parallel edge, other relation
    all_triples[2] = all_triples[0]                                                          # exact duplicate edge
    graph = nx.MultiDiGraph()
    graph.add_weighted_edges_from(all_triples.tolist())                                      # train.py:302
    entities = torch.from_numpy(rng.permutation(n_ids)[:30].astype(np.int64))                # 10 ids have no table row
    ent2idx = make_ent2idx(entities, n_ids - 1)
    edges = blp_b200.graph_edges(graph)
    assert sorted(map(tuple, edges.tolist())) == sorted(map(tuple, all_triples.tolist()))
    # edges(keys=True) would NOT carry the relation (the ADVICE r1 bug): keys are 0, 1, ... per parallel edge
    assert sorted(k for _, _, k in graph.edges(keys=True)) != sorted(all_triples[:, 2].tolist())
    fidx = TripleFilterIndex(edges, ent2idx)
    test = all_triples[rng.permutation(n_edges)[:64]]
    test = test[(ent2idx[test[:, 0]] >= 0).numpy() & (ent2idx[test[:, 1]] >= 0).numpy()]
    hf, tf = fidx.dense_masks(test, len(entities))
    # utils.get_triple_filters, literally, over the nx graph
    rhf = np.zeros((len(test), len(entities)), bool)
    rtf = np.zeros_like(rhf)
    for i, (head, tail, rel) in enumerate(test.tolist()):
        for (h, t, r) in graph.out_edges(head, data='weight'):
            if r == rel and t != tail and ent2idx[t] != -1:
                rtf[i, ent2idx[t]] = True
        for (h, t, r) in graph.in_edges(tail, data='weight'):
            if r == rel and h != head and ent2idx[h] != -1:
                rhf[i, ent2idx[h]] = True
    assert np.array_equal(hf, rhf) and np.array_equal(tf, rtf)
    assert rhf.sum() + rtf.sum() > 0


def test_filter_index_cache_is_keyed_by_ent2idx_content(monkeypatch):
    """lazy._index_for: the device filter index behind the patched utils.get_triple_filters is reused across the batches
    of one evaluation (same ent2idx OBJECT), across evaluations whose freshly built ent2idx has the same CONTENT, and
    rebuilt when the content differs (validation vs test entities over the same graph, train.py:87, 298-302) -- never
    matched by a recycled address."""
    import gc

    import networkx as nx

    from blp_b200 import lazy, utils as butils
    built = []

    class FakeIndex:
        def __init__(self, edges, ent2idx, n, r, dev):
            built.append(ent2idx.clone())
    monkeypatch.setattr(butils, "DeviceFilterIndex", FakeIndex)
    graph = nx.MultiDiGraph()
    graph.add_weighted_edges_from([(0, 1, 2), (1, 2, 0)])
    e_val = torch.arange(5)
    i1, _ = lazy._index_for(graph, e_val, 5, "cpu")
    assert lazy._index_for(graph, e_val, 5, "cpu")[0] is i1 and len(built) == 1          # per-batch calls: same object
    assert lazy._index_for(graph, torch.arange(5), 5, "cpu")[0] is i1 and len(built) == 1  # next evaluation, same content
    e_test = torch.tensor([0, 1, -1, 3, 4])
    i2, _ = lazy._index_for(graph, e_test, 5, "cpu")
    assert i2 is not i1 and len(built) == 2                                                # other entity set: rebuilt
    assert lazy._index_for(graph, e_val, 5, "cpu")[0] is i1 and len(built) == 2          # both stay cached
    e_test[2] = 2                                                                          # in-place change is noticed
    assert lazy._index_for(graph, e_test, 5, "cpu")[0] is i1 and len(built) == 2
    key = (id(graph), 5, "cpu")
    assert key in lazy._INDEX_CACHE
    del graph
    gc.collect()
    assert key not in lazy._INDEX_CACHE                                                    # entries die with the graph


def test_k_values_readback_is_cached_per_tensor_object():
    """ops._kvalues: the reference passes one (device) tensor to every get_metrics call of an evaluation (train.py:72);
    it is read back once per tensor object and version, again after an in-place change or for another tensor."""
    from blp_b200 import ops
    k = torch.tensor([[1, 3, 10]])
    assert ops._kvalues(k)[0] == [1, 3, 10]
    calls = []
    real = torch.Tensor.tolist

    def counting(self):
        calls.append(1)
        return real(self)
    torch.Tensor.tolist = counting
    try:
        assert ops._kvalues(k)[0] == [1, 3, 10] and not calls          # same object, same version: no read-back
        k[0, 1] = 5
        assert ops._kvalues(k)[0] == [1, 5, 10] and len(calls) == 1    # in-place change
        assert ops._kvalues(torch.tensor([1, 10]))[0] == [1, 10] and len(calls) == 2
        assert ops._kvalues([1, 3])[0] == [1, 3] and len(calls) == 2    # plain sequences never touch a tensor
    finally:
        torch.Tensor.tolist = real
    ks, arr = ops._kvalues(())
    assert ks == [] and len(arr) == 1


def test_new_entity_mask_is_built_once_per_set_object():
    """utils._new_entity_mask (behind the patched utils.split_by_new_position): one membership mask per `new_entities`
    set object (train.py passes the same set to every batch), rebuilt when the set grows or another set arrives."""
    from blp_b200 import utils as butils
    dev = torch.device("cpu")
    s = {3, 7, 11}
    m1 = butils._new_entity_mask(s, dev)
    assert m1.dtype == torch.uint8 and m1.tolist() == [0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1]
    assert butils._new_entity_mask(s, dev) is m1
    s.add(1)
    m2 = butils._new_entity_mask(s, dev)
    assert m2 is not m1 and m2[1] == 1
    assert butils._new_entity_mask({3, 7, 11, 1}, dev) is not m2           # an equal but different object is re-read
    assert butils._new_entity_mask(set(), dev).tolist() == [0]
    mask = torch.tensor([True, False, True])
    assert butils._new_entity_mask(mask, dev).tolist() == [1, 0, 1]


def _np_score_fn(model):
    """CPU stand-in for the score kernels (test seam of topk_sweep): the NumPy restatement of models.py:222-248."""
    fn = getattr(np_oracle, model + "_score")

    def score(heads, tails, rels):
        return torch.from_numpy(np.asarray(fn(heads.numpy(), tails.numpy(), rels.numpy()), dtype=np.float32))
    return score


def _topk_inputs(model, n=97, t=11, n_rel=5, seed=3):
    g = torch.Generator().manual_seed(seed)
    ent = torch.randn(n, 16, generator=g)
    ent[40] = ent[7]                                   # tied candidates in different shards: order must be by row
    ent[96] = ent[7]
    rel = torch.randn(n_rel, 16, generator=g) * 0.3
    triples = torch.stack([torch.randint(0, n, (t,), generator=g), torch.randint(0, n, (t,), generator=g),
                           torch.randint(0, n_rel, (t,), generator=g)], dim=1)
    return ent, rel, triples


def _topk_worker(rank, world, port, model, k, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ent, rel, triples = _topk_inputs(model)
        lo, hi = shard_bounds(ent.shape[0], world, rank)
        out = blp_b200.topk_sweep(model, ent[lo:hi].contiguous(), rel, triples, k=k, ent_offset=lo, group=dist.group.WORLD,
                                  chunk=4, score_fn=_np_score_fn(model))
        ret[rank] = (out["scores"].numpy(), out["index"].numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("model", ("transe", "distmult"))
@pytest.mark.parametrize("world,k", ((2, 10), (3, 40)))
def test_topk_sweep_sharded_equals_single_rank_gloo(model, world, k):
    """topk_sweep: per-shard top-k + ONE all-gather + merge gives the single-rank list bit for bit (scores and rows,
    ties by ascending row), also when k exceeds a shard's row count; and the single-rank list is the stable descending
    sort of the full score rows."""
    import torch.multiprocessing as mp
    ent, rel, triples = _topk_inputs(model)
    single = blp_b200.topk_sweep(model, ent, rel, triples, k=k, score_fn=_np_score_fn(model))
    fn = _np_score_fn(model)
    r = rel[triples[:, 2]].unsqueeze(1)
    full = (fn(ent.unsqueeze(0), ent[triples[:, 1]].unsqueeze(1), r), fn(ent[triples[:, 0]].unsqueeze(1), ent.unsqueeze(0), r))
    for role in range(2):
        sv, si = torch.sort(full[role], dim=1, descending=True, stable=True)
        assert torch.equal(single["scores"][role], sv[:, :k]) and torch.equal(single["index"][role], si[:, :k])
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_topk_worker, args=(world, _free_port(), model, k, ret), nprocs=world, join=True)
    for rank in range(world):
        assert np.array_equal(ret[rank][0], single["scores"].numpy()), rank
        assert np.array_equal(ret[rank][1], single["index"].numpy()), rank
    few = blp_b200.topk_sweep(model, ent[:3].contiguous(), rel, triples % 3, k=5, score_fn=_np_score_fn(model))
    assert bool((few["index"][..., 3:] == -1).all()) and bool(torch.isinf(few["scores"][..., 3:]).all())
