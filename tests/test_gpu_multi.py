"""Entity-sharded eval sweep over NCCL: one process per GPU, row-block shards, ONE all-reduce of the
int32 rank counters per sweep.  Integer sums are exact, so every world size must reproduce the single-GPU (and the
reference's golden) numbers bit for bit.  Needs >= 2 GPUs on the box; skipped otherwise."""
import os
import socket

import numpy as np
import pytest
import torch

from conftest import golden

import blp_b200

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _inputs(model):
    g = golden("eval_loop_" + model)
    ent2idx = blp_b200.make_ent2idx(torch.from_numpy(g["entities"]), int(g["n_ids"]) - 1)
    triples = torch.from_numpy(g["triples"])
    rows = torch.stack([ent2idx[triples[:, 0]], ent2idx[triples[:, 1]], triples[:, 2]], dim=1)
    return g, ent2idx, rows


def _worker(rank, world, port, model, mode, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        g, ent2idx, rows = _inputs(model)
        table = torch.from_numpy(g["ent_emb"])
        lo, hi = blp_b200.shard_bounds(table.shape[0], world, rank)
        shard = table[lo:hi].contiguous().to(dev)
        rel = torch.from_numpy(g["rel_weight"]).to(dev)
        didx = blp_b200.DeviceFilterIndex(g["graph_edges"], ent2idx, table.shape[0], rel.shape[0], dev)
        out = blp_b200.rank_sweep(model, shard, rel, rows.to(dev), filter_index=didx, chunk=40, ent_offset=lo,
                                  group=dist.group.WORLD, mode=mode)
        m = blp_b200.finalize(out)
        if rank == world - 1:                      # any rank holds the full result after the all-reduce
            ret.update({k: out[k].cpu().numpy() for k in ("gt", "ge", "gt_f", "ge_f")})
            ret["metrics"] = m
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("model,mode", [("transe", "exact"), ("complex", "exact"), ("distmult", "fast"),
                                        ("complex", "fast_exact"), ("distmult", "fast_exact")])
def test_nccl_sharded_sweep_equals_single_gpu(model, mode, cuda_device):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    g, ent2idx, rows = _inputs(model)
    ent, rel = torch.from_numpy(g["ent_emb"]).to(cuda_device), torch.from_numpy(g["rel_weight"]).to(cuda_device)
    didx = blp_b200.DeviceFilterIndex(g["graph_edges"], ent2idx, ent.shape[0], rel.shape[0], cuda_device)
    single = blp_b200.rank_sweep(model, ent, rel, rows.to(cuda_device), filter_index=didx, mode="exact")
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), model, mode, ret), nprocs=world, join=True)
    if mode in ("exact", "fast_exact"):       # fast_exact: tensor-core sweep + exact refine per shard, the same integers
        for k in ("gt", "ge", "gt_f", "ge_f"):
            assert np.array_equal(ret[k], single[k].cpu().numpy()), k
        want = dict(zip(g["scalar_names"].tolist(), g["scalar_values"].tolist()))
        assert abs(ret["metrics"]["mrr"] - want["test_mrr"]) <= 1e-6
        assert abs(ret["metrics"]["mrr_f"] - want["test_mrr_filt"]) <= 1e-6
    else:
        # tensor-core mode: tolerance-classified (tests/test_gpu_fast.py); across shards the counts still add up to
        # something within the tie band of the exact ranks -- on this tiny table they coincide except for near-ties
        diff = np.abs(ret["gt"].astype(np.int64) - single["gt"].cpu().numpy().astype(np.int64))
        assert diff.max() <= 2 and (diff > 0).mean() < 0.05


@pytest.mark.parametrize("model,loss", [("transe", "margin"), ("complex", "nll")])
def test_dataparallel_threads_compute_loss_concurrently(model, loss, cuda_device):
    """The reference's only multi-GPU mechanism (train.py:329-330): torch.nn.DataParallel runs forward() from one Python
    thread per device, concurrently, each on its sub-batch with LOCAL in-batch negatives (data.py:289-298, repeats =
    number of devices).  The C ABI must be re-entrant (thread-local error / launch state, per-(device, stream)
    workspaces, device taken from the tensors): per-replica losses and the reduced gradients must equal the oracle's."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from oracle import c_oracle
    n_dev, b_per, k, d, n_ent, n_rel = 2, 24, 40, 128, 300, 9
    torch.manual_seed(5)
    m = blp_b200.TransductiveLinkPrediction(d, model, loss, n_ent, n_rel, 1e-3 if model == "complex" else 0)
    dp = torch.nn.DataParallel(m, device_ids=list(range(n_dev))).to(torch.device("cuda", 0))
    g = torch.Generator().manual_seed(6)
    pos_pairs = torch.randint(0, n_ent, (n_dev * b_per, 2), generator=g)
    rels = torch.randint(0, n_rel, (n_dev * b_per, 1), generator=g)
    # data.py:289-298: indices local to each device's sub-batch of b_per pairs
    neg_idx = torch.cat([blp_b200.get_negative_sampling_indices(b_per, k, device=cuda_device, seed=7 + i).contiguous().cpu()
                         for i in range(n_dev)])
    for it in range(3):                                  # repeated steps: replicas are rebuilt, threads re-enter
        dp.zero_grad()
        per_replica = dp(pos_pairs.to(cuda_device), rels.to(cuda_device), neg_idx.to(cuda_device))
        assert per_replica.shape == (n_dev,)
        per_replica.mean().backward()                    # train.py:343-346
    ent_w, rel_w = m.ent_emb.weight.detach().cpu(), m.rel_emb.weight.detach().cpu()
    want_grad_rel = np.zeros((n_rel, d), np.float64)
    for i in range(n_dev):
        sl = slice(i * b_per, (i + 1) * b_per)
        embs = ent_w[pos_pairs[sl]]
        if model == "transe":
            embs = torch.nn.functional.normalize(embs, dim=-1)
        co = c_oracle.train_loss(model, loss, embs.numpy(), rel_w[rels[sl, 0]].numpy(), neg_idx[sl].numpy(),
                                 float(m.regularizer))
        assert abs(float(per_replica[i]) - float(co["loss"])) <= 1e-5 * abs(float(co["loss"])), i
        np.add.at(want_grad_rel, rels[sl, 0].numpy(), co["grad_rel"].astype(np.float64) / n_dev)
    got = m.rel_emb.weight.grad.cpu().numpy()
    assert np.abs(got - want_grad_rel).max() <= 2e-5 * np.abs(want_grad_rel).max()


def _topk_worker(rank, world, port, model, k, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        g, ent2idx, rows = _inputs(model)
        table = torch.from_numpy(g["ent_emb"])
        lo, hi = blp_b200.shard_bounds(table.shape[0], world, rank)
        out = blp_b200.topk_sweep(model, table[lo:hi].contiguous().to(dev), torch.from_numpy(g["rel_weight"]).to(dev), rows.to(dev),
                                  k=k, ent_offset=lo, group=dist.group.WORLD, chunk=16)
        if rank == 0:
            ret.update(scores=out["scores"].cpu().numpy(), index=out["index"].cpu().numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("model", ("transe", "complex"))
def test_nccl_sharded_topk_equals_single_gpu(model, cuda_device):
    """topk_sweep over row shards: per-shard top-k, ONE NCCL all-gather of the packed (score bits, row) pairs, merge --
    the same scores and rows as the single-GPU list, bit for bit."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    g, ent2idx, rows = _inputs(model)
    ent, rel = torch.from_numpy(g["ent_emb"]).to(cuda_device), torch.from_numpy(g["rel_weight"]).to(cuda_device)
    single = blp_b200.topk_sweep(model, ent, rel, rows.to(cuda_device), k=10)
    ret = mp.Manager().dict()
    mp.spawn(_topk_worker, args=(world, _free_port(), model, 10, ret), nprocs=world, join=True)
    assert np.array_equal(ret["scores"], single["scores"].cpu().numpy())
    assert np.array_equal(ret["index"], single["index"].cpu().numpy())
