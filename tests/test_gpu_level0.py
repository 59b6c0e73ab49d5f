"""Integration levels side by side on an FB15k-237-shaped evaluation:

  Level 0   the reference's OWN, byte-identical train.eval_link_prediction (oracle/_ref) after blp_b200.patch(models,
            utils): the per-batch loop of train.py:128-171 runs on score-matrix handles (one fused launch per batch)
  Level 0d  the same with lazy_scores=False: the reference's statements on real (2B, N) matrices produced by the exact
            score kernels + blp_rank_counts
  Level 1   blp_b200.rank_sweep over the whole set + DeviceFilterIndex + finalize (one launch per sweep)

All three must report the same raw / filtered MRR and hits@k (integer ranks are bit-exact; the scalars are sums of the
same reciprocal ranks in different orders, so 1e-6).  With BLP_LEVEL0_OUT=<file> the wall-clock time of each level is
written there (tools/gpu_level0.sh -> profiles/): what a user of the reference gets without touching train.py, and
what the 25-line replacement of the rank block adds.
"""
import logging
import os
import time
import types

import numpy as np
import pytest
import torch

import blp_b200
from blp_b200 import lazy
from oracle import ref_loader

pytestmark = pytest.mark.gpu

N_ENT, N_REL, DIM = 14541, 237, 128
N_EDGES = 60000             # filtering graph (train + valid + test triples of the synthetic graph)
EVAL_BATCH = 64             # the reference's eval_batch_size


class _Text:
    def get_entity_description(self, ents):
        tok = ents.reshape(-1, 1).repeat(1, 4)
        return tok, torch.ones_like(tok, dtype=torch.float), torch.full((tok.shape[0],), 4)


class _Loader:
    def __init__(self, triples, bs, rel_categories):
        self.batches = list(torch.split(triples, bs)) if triples.shape[0] else []
        self.dataset = types.SimpleNamespace(rel_categories=rel_categories, has_rel_categories=True)

    def __iter__(self):
        return iter(self.batches)

    def __len__(self):
        return len(self.batches)


class _Run:
    def __init__(self):
        self.scalars = {}

    def log_scalar(self, name, value, step=None):
        self.scalars[name] = float(value)


def _setup(model_name, n_test, seed=5):
    import networkx as nx
    g = torch.Generator().manual_seed(seed)
    edges = torch.stack([torch.randint(0, N_ENT, (N_EDGES,), generator=g), torch.randint(0, N_ENT, (N_EDGES,), generator=g),
                         torch.randint(0, N_REL, (N_EDGES,), generator=g)], dim=1)
    triples = edges[:n_test].clone()                       # test triples are edges of the filtering graph (train.py:298-302)
    graph = nx.MultiDiGraph()
    graph.add_weighted_edges_from(edges.tolist())          # as train.py:298-302: the relation id is the edge weight
    graph.add_nodes_from(range(N_ENT))
    table = torch.randn(N_ENT, DIM, generator=g)
    rel = (torch.rand(N_REL, DIM, generator=g) * 2 - 1) * 0.128
    rel_categories = torch.randint(0, 4, (N_REL,), generator=g)
    return dict(graph=graph, triples=triples, table=table, rel=rel, rel_categories=rel_categories,
                entities=torch.arange(N_ENT), new_entities=set(range(0, N_ENT, 3)))


def _sync(device):
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize()


def _run_reference_loop(model_name, s, device, lazy_scores, repeats, with_breakdowns=True, patched=True):
    ref = ref_loader.load(("models", "utils", "train"))
    ref_models, ref_utils, ref_train = ref["models"], ref["utils"], ref["train"]
    saved = {m: dict(vars(m)) for m in (ref_models, ref_utils)}
    saved_cl = ref_models.LinkPrediction.compute_loss
    saved_dev = ref_train.device
    try:
        if patched:
            blp_b200.patch(ref_models, ref_utils, lazy_scores=lazy_scores)
        ref_train.device = device

        class TableEncoder(ref_models.InductiveLinkPrediction):
            def __init__(self, dim, rel_model, num_entities, num_relations):
                super().__init__(dim, rel_model, "margin", num_relations, 0)
                self.table = torch.nn.Embedding(num_entities, dim)

            def _encode_entity(self, text_tok, text_mask):
                return self.table(text_tok[:, 0])

        m = TableEncoder(DIM, model_name, N_ENT, N_REL)
        with torch.no_grad():
            m.table.weight.copy_(s["table"])
            m.rel_emb.weight.copy_(s["rel"])
        m = m.to(device)
        loader = _Loader(s["triples"], EVAL_BATCH, s["rel_categories"])
        loader.dataset.has_rel_categories = with_breakdowns          # train.py:183: the by-category Python loop
        new_entities = s["new_entities"] if with_breakdowns else None   # train.py:174: the by-position Python loop
        wrapped = types.SimpleNamespace(module=m) if torch.device(device).type == "cuda" else m   # train.py:79-80: `model = model.module` on a GPU
        times = []
        if os.environ.get("BLP_LEVEL0_PROFILE") and patched and lazy_scores and not with_breakdowns:
            # host profile of the untouched loop (where do the microseconds per batch go?)
            import cProfile
            import io
            import pstats
            ref_train.eval_link_prediction(wrapped, loader, _Text(), s["entities"], 0, 4096, _Run(), logging.getLogger("level0"),
                                           prefix="test", filtering_graph=s["graph"], new_entities=new_entities)
            pr = cProfile.Profile()
            pr.enable()
            ref_train.eval_link_prediction(wrapped, loader, _Text(), s["entities"], 0, 4096, _Run(), logging.getLogger("level0"),
                                           prefix="test", filtering_graph=s["graph"], new_entities=new_entities)
            pr.disable()
            buf = io.StringIO()
            st = pstats.Stats(pr, stream=buf)
            st.sort_stats("cumulative").print_stats(45)
            st.sort_stats("tottime").print_stats(30)
            with open(os.environ["BLP_LEVEL0_PROFILE"], "a") as f:
                f.write(f"==== {model_name}: cProfile of one eval_link_prediction call, {len(loader)} batches\n" + buf.getvalue())
        for _ in range(repeats):
            run = _Run()
            _sync(device)
            t0 = time.perf_counter()
            mrr, ent_emb = ref_train.eval_link_prediction(
                wrapped, loader, _Text(), s["entities"], 0, 4096, run, logging.getLogger("level0"), prefix="test",
                filtering_graph=s["graph"], new_entities=new_entities, return_embeddings=True)
            _sync(device)
            times.append(time.perf_counter() - t0)
        # the encoder pass alone (train.py:95-123): the same call with no test batches
        empty = _Loader(s["triples"][:0], EVAL_BATCH, s["rel_categories"])
        empty.dataset.has_rel_categories = False
        enc = []
        for _ in range(2):
            _sync(device)
            t0 = time.perf_counter()
            try:
                ref_train.eval_link_prediction(wrapped, empty, _Text(), s["entities"], 0, 4096, _Run(), logging.getLogger("level0"),
                                               prefix="test", filtering_graph=s["graph"], new_entities=None)
            except ZeroDivisionError:                        # train.py:196 divides by the number of predictions
                pass
            _sync(device)
            enc.append(time.perf_counter() - t0)
        return run.scalars, ent_emb.squeeze(0), m.rel_emb.weight.detach(), min(times), min(enc)
    finally:
        lazy.enable(False)
        ref_train.device = saved_dev
        for mod, attrs in saved.items():
            for k, v in attrs.items():
                setattr(mod, k, v)
        ref_models.LinkPrediction.compute_loss = saved_cl


@pytest.mark.skipif(ref_loader.available() is None, reason="reference modules not built (python oracle/build_ref.py)")
@pytest.mark.parametrize("model_name", ("transe", "distmult"))
def test_integration_levels_agree(model_name, cuda_device):
    timing = bool(os.environ.get("BLP_LEVEL0_OUT"))
    n_test = 2048 if timing else 512
    s = _setup(model_name, n_test)
    dev = cuda_device
    sc0, ent_emb, rel_w, t0, enc0 = _run_reference_loop(model_name, s, dev, True, 3 if timing else 1)
    sc0d, _, _, t0d, _ = _run_reference_loop(model_name, s, dev, False, 2 if timing else 1)
    if timing:      # the rank block alone: without the reference's per-triple Python loops (train.py:173-188)
        sc0p, _, _, t0p, _ = _run_reference_loop(model_name, s, dev, True, 3, with_breakdowns=False)
        _, _, _, t0dp, _ = _run_reference_loop(model_name, s, dev, False, 2, with_breakdowns=False)
        # the unmodified, unpatched reference on this box's host cores (first 512 test triples)
        s_cpu = dict(s, triples=s["triples"][:512])
        sc_cpu, _, _, t_cpu, enc_cpu = _run_reference_loop(model_name, s_cpu, torch.device("cpu"), False, 1, with_breakdowns=False,
                                                           patched=False)

    # ---- Level 1: the rank block replaced by one fused sweep
    rows = s["triples"].to(dev)                            # ids are rows here (ent2idx is the identity)
    times1, build = [], None
    for _ in range(3 if timing else 1):
        torch.cuda.synchronize()
        t = time.perf_counter()
        fidx = blp_b200.DeviceFilterIndex(blp_b200.graph_edges(s["graph"]), None, N_ENT, N_REL, dev)
        torch.cuda.synchronize()
        build = time.perf_counter() - t
        t = time.perf_counter()
        out = blp_b200.rank_sweep(model_name, ent_emb, rel_w, rows, filter_index=fidx)
        m1 = blp_b200.finalize(out)
        torch.cuda.synchronize()
        times1.append(time.perf_counter() - t)
        t = time.perf_counter()
        bd = blp_b200.breakdowns(out, s["triples"], new_entities=s["new_entities"], rel_categories=s["rel_categories"])
        torch.cuda.synchronize()
        t_bd = time.perf_counter() - t
    modes = {}
    if model_name != "transe":
        t = None
        for _ in range(3 if timing else 1):
            torch.cuda.synchronize()
            t = time.perf_counter()
            outx = blp_b200.rank_sweep(model_name, ent_emb, rel_w, rows, filter_index=fidx, mode="fast_exact")
            mx = blp_b200.finalize(outx)
            torch.cuda.synchronize()
            t = time.perf_counter() - t
        assert "refine_overflow" not in outx
        for k in ("gt", "ge", "gt_f", "ge_f"):
            assert torch.equal(out[k], outx[k]), k
        modes["fast_exact"] = t

    def check(sc, tag):
        assert abs(sc["test_mrr"] - m1["mrr"]) <= 1e-6, tag
        assert abs(sc["test_mrr_filt"] - m1["mrr_f"]) <= 1e-6, tag
        for j, k in enumerate((1, 3, 10)):
            assert abs(sc[f"test_hits@{k}"] - m1["hits_at_k"][j]) <= 1e-9, tag
            assert abs(sc[f"test_hits@{k}_filt"] - m1["hits_at_k_f"][j]) <= 1e-9, tag
    check(sc0, "level 0 (score-matrix handles)")
    check(sc0d, "level 0 (materialised matrices)")
    assert m1["mrr_f"] >= m1["mrr"] > 0.0
    # by-position breakdown (train.py:215-225): the reference's scalars vs the device accumulators of Level 1
    pos = (bd["mrr_by_position"] / bd["mrr_pos_counts"].clamp(min=1.0)).tolist()
    for j, name in enumerate(("test_mrr_filt_both_new", "test_mrr_filt_head_new", "test_mrr_filt_tail_new")):
        assert abs(sc0[name] - pos[j]) <= 2e-6, name

    if timing:
        nb = n_test // EVAL_BATCH
        with open(os.environ["BLP_LEVEL0_OUT"], "a") as f:
            f.write(f"{model_name}: {n_test} test triples in {nb} batches of {EVAL_BATCH}, {N_ENT} entities, d = {DIM}, filtering graph of "
                    f"{N_EDGES} edges; raw + filtered ranks\n")
            f.write(f"  encoder pass of eval_link_prediction (train.py:95-123, table encoder, batches of 4096): {enc0 * 1e3:8.2f} ms\n")
            f.write(f"  Level 0  unmodified train.eval_link_prediction + patch():            {t0 * 1e3:8.2f} ms total, "
                    f"{(t0 - enc0) / nb * 1e6:7.1f} us per batch after the encoder pass\n")
            f.write(f"  Level 0d the same with lazy_scores=False (real (2B, N) matrices):     {t0d * 1e3:8.2f} ms total, "
                    f"{(t0d - enc0) / nb * 1e6:7.1f} us per batch\n")
            f.write(f"  Level 0  without the per-triple Python loops of train.py:173-188:     {t0p * 1e3:8.2f} ms total, "
                    f"{(t0p - enc0) / nb * 1e6:7.1f} us per batch\n")
            f.write(f"  Level 0d without them:                                                {t0dp * 1e3:8.2f} ms total, "
                    f"{(t0dp - enc0) / nb * 1e6:7.1f} us per batch\n")
            f.write(f"  Level 1  DeviceFilterIndex build (once per evaluation):               {build * 1e3:8.2f} ms\n")
            f.write(f"  Level 1  rank_sweep + finalize (whole set, one launch + correction):  {min(times1) * 1e3:8.2f} ms total, "
                    f"{min(times1) / nb * 1e6:7.1f} us per 64 triples\n")
            f.write(f"  Level 1  breakdowns (by position / by category, device accumulators): {t_bd * 1e3:8.2f} ms\n")
            f.write(f"  CPU      the unpatched reference on the host ({torch.get_num_threads()} threads), 512 triples, no Python breakdown loops: "
                    f"{(t_cpu - enc_cpu) / 8 * 1e6:9.1f} us per batch\n")
            for k, v in modes.items():
                f.write(f"  Level 1  rank_sweep(mode='{k}') + finalize:                   {v * 1e3:8.2f} ms total\n")
