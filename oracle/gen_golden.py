"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Run in the build container (the only place /root/reference exists):

    python oracle/gen_golden.py            # writes tests/golden/*.npz

What it pins (the reference ships no tests or fixtures -- SURVEY.md section 4):
  * models.py:222-248 score functions on eval-shaped and train-shaped
    broadcasts, utils.py:86-111 get_metrics, train.py:159-167 filtered
    re-rank                                               -> eval_<model>_d<D>.npz
  * models.py:51-70 compute_loss forward + autograd backward with neg_idx from
    the reference's own sampler (data.py:35-81), incl. its non-contiguous
    strides                                               -> train_<model>_<loss>.npz
  * train.py:57-243 eval_link_prediction end to end on a synthetic inductive
    dataset with a filtering graph (raw + filtered MRR / hits@k, by-position
    and by-category breakdowns)                           -> eval_loop_<model>.npz
  * Appendix-B known-answer cases                         -> kat.npz
While generating, it asserts that both oracle restatements (blp_oracle.c and
np_oracle.py) are BIT-EQUAL to the reference on every score and integer rank.
"""
import logging
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("BLP_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def import_reference():
    """Import reference models/utils/data/train with nltk + sacred stubbed (neither is installed)."""
    nltk = types.ModuleType("nltk")
    nltk.download = lambda *a, **k: True
    nltk.word_tokenize = lambda s: s.split()
    corpus = types.ModuleType("nltk.corpus")
    corpus.stopwords = types.SimpleNamespace(words=lambda lang: [])
    nltk.corpus = corpus
    sys.modules.update({"nltk": nltk, "nltk.corpus": corpus})

    sacred = types.ModuleType("sacred")

    class Experiment:
        def __init__(self, *a, **k):
            self.observers = []
            self.logger = None

        def _ident(self, f):
            return f
        config = capture = command = automain = main = _ident

        def run_commandline(self, *a, **k):
            return None
    sacred.Experiment = Experiment
    run_mod = types.ModuleType("sacred.run")
    run_mod.Run = object
    obs = types.ModuleType("sacred.observers")
    obs.MongoObserver = object
    sys.modules.update({"sacred": sacred, "sacred.run": run_mod, "sacred.observers": obs})

    sys.path.insert(0, REF)
    import models  # noqa
    import utils  # noqa
    import data  # noqa
    import train  # noqa
    return models, utils, data, train


ref_models, ref_utils, ref_data, ref_train = import_reference()

from oracle import c_oracle, np_oracle  # noqa: E402

REF_SCORE = {"transe": ref_models.transe_score, "distmult": ref_models.distmult_score,
             "complex": ref_models.complex_score, "simple": ref_models.simple_score}
REF_LOSS = {"margin": ref_models.margin_loss, "nll": ref_models.nll_loss}


def make_table(n, d, normalize, seed):
    g = torch.Generator().manual_seed(seed)
    ent = torch.randn(n, d, generator=g)
    if normalize:  # models.py:40-41 (TransE tables are L2-normalised rows)
        ent = torch.nn.functional.normalize(ent, dim=-1)
    return ent


def make_rel(r, d, seed):
    g = torch.Generator().manual_seed(seed)
    a = (6.0 / (r + d)) ** 0.5  # xavier_uniform_, models.py:29
    return (torch.rand(r, d, generator=g) * 2 - 1) * a


def assert_bits(name, a, b):
    a, b = np.asarray(a), np.asarray(b)
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        raise AssertionError(f"{name}: {len(bad)} of {a.size} differ, first at {bad[0]}: {a[tuple(bad[0])]} vs {b[tuple(bad[0])]}")


def gen_eval(model, n, d, b, seed, tie_rows=True):
    """One eval batch exactly as train.py:141-171 does it (incl. filtered re-rank)."""
    g = torch.Generator().manual_seed(seed + 100)
    ent = make_table(n, d, model == "transe", seed)
    if tie_rows:
        # duplicate a few rows so exact score ties (gt != ge - 1) are exercised
        ent[n - 1] = ent[3]
        ent[n - 2] = ent[5]
    rel = make_rel(11, d, seed + 1)
    heads = torch.randint(0, n, (b, 1), generator=g)
    tails = torch.randint(0, n, (b, 1), generator=g)
    heads[0, 0] = 3      # a true entity that has an exact duplicate in the table
    tails[1, 0] = 5
    rels = torch.randint(0, 11, (b, 1), generator=g)
    ent_emb = ent.unsqueeze(0)
    head_embs = ent_emb.squeeze()[heads]
    tail_embs = ent_emb.squeeze()[tails]
    rel_embs = rel[rels]
    fn = REF_SCORE[model]
    hp = fn(ent_emb, tail_embs, rel_embs)
    tp = fn(head_embs, ent_emb, rel_embs)
    pred = torch.cat((hp, tp))
    true = torch.cat((heads, tails))
    k_values = torch.tensor([[1, 3, 10]])
    recip, hits = ref_utils.get_metrics(pred, true, k_values)
    st = pred.gather(1, true)
    gt = (pred > st).sum(1)
    ge = (pred >= st).sum(1)
    # random filter mask, never covering the true entity (utils.py:71,78)
    mask = torch.rand(2 * b, n, generator=g) < 0.05
    mask[torch.arange(2 * b), true.squeeze()] = False
    mask[2] = False  # an empty filter row
    pred_f = pred.clone()
    pred_f[mask] = pred_f.min() - 1.0
    recip_f, hits_f = ref_utils.get_metrics(pred_f, true, k_values)
    st_f = pred_f.gather(1, true)
    gt_f = (pred_f > st_f).sum(1)
    ge_f = (pred_f >= st_f).sum(1)

    # --- the oracle must reproduce the reference bit for bit ---
    rel_rows = rel_embs.squeeze(1).numpy()
    h_rows, t_rows = head_embs.squeeze(1).numpy(), tail_embs.squeeze(1).numpy()
    indptr = np.concatenate([[0], np.cumsum(mask.sum(1).numpy())]).astype(np.int64)
    idx = np.nonzero(mask.numpy())[1].astype(np.int64)
    co = c_oracle.eval_rank(model, ent.numpy(), h_rows, t_rows, rel_rows, heads.squeeze(1).numpy(),
                            tails.squeeze(1).numpy(), indptr, idx, want_scores=True)
    assert_bits(f"{model} d={d} C scores", pred.numpy(), co["scores"])
    assert_bits("gt", gt.numpy(), co["gt"]); assert_bits("ge", ge.numpy(), co["ge"])
    assert_bits("gt_f", gt_f.numpy(), co["gt_f"]); assert_bits("ge_f", ge_f.numpy(), co["ge_f"])
    r2, h2 = c_oracle.metrics_from_counts(co["gt"], co["ge"], [1, 3, 10])
    assert_bits("recip", recip.numpy(), r2); assert_bits("hits", hits.numpy(), h2)
    r2, h2 = c_oracle.metrics_from_counts(co["gt_f"], co["ge_f"], [1, 3, 10])
    assert_bits("recip_f", recip_f.numpy(), r2); assert_bits("hits_f", hits_f.numpy(), h2)
    no = np_oracle.eval_rank_batch(model, ent.numpy(), heads.numpy(), tails.numpy(), rel_rows, filter_mask=mask.numpy())
    assert_bits(f"{model} d={d} numpy scores", pred.numpy(), no["pred"])
    assert_bits("np gt_f", gt_f.numpy(), no["gt_f"]); assert_bits("np recip_f", recip_f.numpy(), no["recip_f"])

    np.savez_compressed(
        os.path.join(OUT, f"eval_{model}_d{d}.npz"), ent=ent.numpy(), rel=rel.numpy(),
        heads=heads.numpy(), tails=tails.numpy(), rels=rels.numpy(),
        pred=pred.numpy(), gt=gt.numpy(), ge=ge.numpy(), recip=recip.numpy(), hits=hits.numpy(),
        filter_mask=np.packbits(mask.numpy(), axis=1), gt_f=gt_f.numpy(), ge_f=ge_f.numpy(),
        recip_f=recip_f.numpy(), hits_f=hits_f.numpy())
    ties = int((ge - gt > 1).sum())
    print(f"eval   {model:8s} d={d:4d} n={n} b={b}: ok (queries with exact ties: {ties})")


def gen_train(model, loss, b, k, d, regularizer, seed):
    """models.py:51-70 through the reference's own LinkPrediction subclass and sampler."""
    torch.manual_seed(seed)
    m = ref_models.TransductiveLinkPrediction(d, model, loss, num_entities=50, num_relations=7, regularizer=regularizer)
    pos_pairs = torch.randint(0, 50, (b, 2))
    rels = torch.randint(0, 7, (b, 1))
    neg_idx = ref_data.get_negative_sampling_indices(b, k)  # non-contiguous view, data.py:78-79
    assert not neg_idx.is_contiguous()
    ent_embs = m.encode(pos_pairs).detach().clone().requires_grad_(True)   # (B,2,D); TransE rows normalised
    out = m.compute_loss(ent_embs, rels, neg_idx)
    out.backward()
    with torch.no_grad():
        rel_rows = m.rel_emb(rels)
        heads, tails = torch.chunk(ent_embs, 2, dim=1)
        pos = m.score_fn(heads, tails, rel_rows)
        neg_embs = ent_embs.view(b * 2, -1)[neg_idx]
        nh, nt = torch.chunk(neg_embs, 2, dim=2)
        neg = m.score_fn(nh.squeeze(), nt.squeeze(), rel_rows)
    grad_rel_rows = torch.zeros(b, d).index_add_(0, torch.arange(b), m.rel_emb.weight.grad[rels.squeeze(1)])  # per-row view (for info)
    co = c_oracle.train_loss(model, loss, ent_embs.detach().numpy(), rel_rows.squeeze(1).numpy(),
                             neg_idx.numpy(), regularizer)
    assert_bits(f"{model} pos", pos.squeeze(1).numpy(), co["pos_scores"])
    assert_bits(f"{model} neg", neg.numpy(), co["neg_scores"])
    assert abs(float(co["loss"]) - float(out)) <= 1e-5 * abs(float(out)), (co["loss"], float(out))
    ge = ent_embs.grad.numpy()
    scale = max(1e-30, np.abs(ge).max())
    assert np.abs(co["grad_ent"] - ge).max() <= 1e-5 * scale, np.abs(co["grad_ent"] - ge).max() / scale
    # rel grads: oracle returns per-batch-row grads; scatter-add them like embedding backward
    gw = np.zeros((7, d), np.float64)
    np.add.at(gw, rels.squeeze(1).numpy(), co["grad_rel"].astype(np.float64))
    gwr = m.rel_emb.weight.grad.numpy()
    assert np.abs(gw - gwr).max() <= 1e-5 * max(1e-30, np.abs(gwr).max())
    nl, npos, nneg = np_oracle.compute_loss(model, loss, ent_embs.detach().numpy(), rel_rows.squeeze(1).numpy(),
                                            neg_idx.numpy(), regularizer)
    assert_bits("np pos", pos.numpy(), npos); assert_bits("np neg", neg.numpy(), nneg)
    assert abs(float(nl) - float(out)) <= 1e-5 * abs(float(out))
    del grad_rel_rows
    np.savez_compressed(
        os.path.join(OUT, f"train_{model}_{loss}.npz"), ent_embs=ent_embs.detach().numpy(),
        rel_weight=m.rel_emb.weight.detach().numpy(), rels=rels.numpy(),
        neg_idx=np.ascontiguousarray(neg_idx.numpy()), neg_idx_strides=np.array(neg_idx.stride()),
        regularizer=np.float32(regularizer), loss=np.float32(out.item()),
        pos_scores=pos.numpy(), neg_scores=neg.numpy(), grad_ent=ge, grad_rel_weight=gwr)
    print(f"train  {model:8s} {loss:6s} b={b} k={k} d={d} reg={regularizer}: ok loss={out.item():.6f}")


class _TableEncoder(ref_models.InductiveLinkPrediction):
    """Minimal inductive model for the end-to-end loop: token 0 carries the entity id."""
    def __init__(self, dim, rel_model, num_entities, num_relations):
        super().__init__(dim, rel_model, "margin", num_relations, 0)
        self.table = torch.nn.Embedding(num_entities, dim)

    def _encode_entity(self, text_tok, text_mask):
        return self.table(text_tok[:, 0])


class _Text:
    def get_entity_description(self, ents):
        tok = ents.reshape(-1, 1).repeat(1, 4)
        return tok, torch.ones_like(tok, dtype=torch.float), torch.full((tok.shape[0],), 4)


class _Loader:
    def __init__(self, triples, bs, rel_categories):
        self.batches = list(torch.split(triples, bs))
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


def gen_eval_loop(model, seed):
    """train.py:57-243 eval_link_prediction end to end (CPU, device == cpu in this container)."""
    import networkx as nx
    torch.manual_seed(seed)
    n_ids, n_rel, d, bs = 90, 5, 128, 16
    ids = torch.randperm(n_ids)
    entities = ids[:70].clone()                     # candidate set of this split (train.py:82-93)
    m = _TableEncoder(d, model, n_ids, n_rel)
    torch.nn.init.normal_(m.table.weight)
    triples = torch.stack([entities[torch.randint(0, 70, (96,))], entities[torch.randint(0, 70, (96,))],
                           torch.randint(0, n_rel, (96,))], dim=1)
    extra = torch.stack([ids[torch.randint(0, n_ids, (600,))], ids[torch.randint(0, n_ids, (600,))],
                         torch.randint(0, n_rel, (600,))], dim=1)
    graph = nx.MultiDiGraph()
    for h, t, r in torch.cat([triples, extra]).tolist():
        graph.add_edge(h, t, weight=r)
    graph.add_nodes_from(range(n_ids))
    rel_categories = torch.randint(0, 4, (n_rel,))
    new_entities = set(entities[:25].tolist())
    loader = _Loader(triples, bs, rel_categories)
    run = _Run()
    log = logging.getLogger("gen")
    mrr, ent_emb = ref_train.eval_link_prediction(
        m, loader, _Text(), entities, 0, 32, run, log, prefix="test", filtering_graph=graph,
        new_entities=new_entities, return_embeddings=True)
    edges = np.array([(h, t, w) for h, t, w in graph.edges(data="weight")], np.int64)
    np.savez_compressed(
        os.path.join(OUT, f"eval_loop_{model}.npz"), table_weight=m.table.weight.detach().numpy(),
        rel_weight=m.rel_emb.weight.detach().numpy(), entities=entities.numpy(), triples=triples.numpy(),
        graph_edges=edges, n_ids=n_ids, rel_categories=rel_categories.numpy(),
        new_entities=np.array(sorted(new_entities)), batch_size=bs, ent_emb=ent_emb.squeeze(0).numpy(),
        scalar_names=np.array(sorted(run.scalars)), scalar_values=np.array([run.scalars[k] for k in sorted(run.scalars)]))
    print(f"loop   {model:8s}: mrr={mrr:.6f} " + " ".join(f"{k}={v:.4f}" for k, v in sorted(run.scalars.items())[:4]))


def gen_kat():
    """SURVEY.md Appendix B known-answer cases, regenerated from the reference."""
    pos = torch.tensor([[1.0], [1.0]], requires_grad=True)
    neg = torch.tensor([[0.0, -1.0, 2.0], [0.5, 0.0, 0.0]], requires_grad=True)
    ml = ref_models.margin_loss(pos, neg); ml.backward()
    p2 = torch.tensor([[0.3]], requires_grad=True); n2 = torch.tensor([[0.1, -0.2]], requires_grad=True)
    nl = ref_models.nll_loss(p2, n2); nl.backward()
    h = torch.tensor([[0.0, 1.0, -2.0]], requires_grad=True)
    t = torch.tensor([[0.0, 0.5, 1.0]]); r = torch.zeros(1, 3)
    ts = ref_models.transe_score(h, t, r); ts.sum().backward()
    e2i = ref_utils.make_ent2idx(torch.tensor([4, 5, 0]), 5)
    ni = ref_data.get_negative_sampling_indices(4, 3)
    np.savez_compressed(
        os.path.join(OUT, "kat.npz"), margin=ml.item(), margin_dpos=pos.grad.numpy(), margin_dneg=neg.grad.numpy(),
        nll=nl.item(), nll_dpos=p2.grad.numpy(), nll_dneg=n2.grad.numpy(), transe=ts.detach().numpy(),
        transe_dh=h.grad.numpy(), ent2idx=e2i.numpy(), neg_idx_shape=np.array(ni.shape), neg_idx_stride=np.array(ni.stride()))
    print("kat    ok: margin=%.7f nll=%.7f" % (ml.item(), nl.item()))


def check_sum_orders():
    g = torch.Generator().manual_seed(0)
    for L in list(range(8, 200)) + [256, 300, 384, 512, 513, 640, 768, 1000, 1024, 2048, 4096]:
        x = torch.randn(5, 9, L, generator=g)
        assert_bits(f"aten_sum L={L}", torch.sum(x, dim=-1).numpy(), np_oracle.aten_sum_lastdim(x.numpy()))
        assert np.float32(torch.sum(x[0, 0]).item()) is not None
        assert_bits(f"c aten_sum L={L}", torch.sum(x[1], dim=-1).numpy(), np.array([c_oracle.aten_sum(v) for v in x[1].numpy()]))
        assert_bits(f"norm1 L={L}", torch.norm(x, dim=-1, p=1).numpy(), np_oracle.seq_sum_lastdim(np.abs(x.numpy())))
    print("orders ok: oracle sum / L1 orders are bit-equal to torch", torch.__version__)


def gen_normalize():
    """models.py:38-43: LinkPrediction.encode -> F.normalize for TransE, through the reference's own class."""
    out = {}
    for d in (128, 300, 768, 100):
        torch.manual_seed(60 + d)
        m = ref_models.TransductiveLinkPrediction(d, "transe", "margin", 40, 3, 0)
        torch.nn.init.normal_(m.ent_emb.weight, std=0.7)
        with torch.no_grad():
            m.ent_emb.weight[3] = 0.0                      # zero row: clamp_min(eps) path
            m.ent_emb.weight[5] *= 1e-20                   # tiny norm
        ents = torch.arange(40)
        y = m.encode(ents).detach()
        x = m.ent_emb.weight.detach()
        assert_bits(f"normalize d={d}", y.numpy(), np_oracle.l2_normalize_rows(x.numpy()))
        out[f"x_{d}"] = x.numpy()
        out[f"y_{d}"] = y.numpy()
    np.savez_compressed(os.path.join(OUT, "normalize.npz"), **out)
    print("normalize ok: oracle l2_normalize_rows is bit-equal to the reference's encode()")


def main():
    os.makedirs(OUT, exist_ok=True)
    if "--normalize-only" in sys.argv:
        gen_normalize()
        return
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    check_sum_orders()
    gen_kat()
    for model in ("transe", "distmult", "complex", "simple"):
        gen_eval(model, n=257, d=128, b=8, seed=10)
    gen_eval("transe", n=131, d=300, b=4, seed=20)     # glove-bow width (utils.py:17-19)
    gen_eval("transe", n=67, d=768, b=3, seed=21)      # bert-bow width (scripts/test-umls.sh)
    gen_eval("distmult", n=99, d=300, b=3, seed=22)    # generic-D bilinear (remainder vectors in the sum)
    gen_eval("complex", n=99, d=200, b=3, seed=23)     # L=100: scalar tail in the sum
    gen_eval("simple", n=65, d=1600, b=2, seed=24)     # L=800: cascade level in the sum
    for i, model in enumerate(("transe", "distmult", "complex", "simple")):
        gen_train(model, "margin", b=8, k=16, d=128, regularizer=0.0 if model != "complex" else 1e-3, seed=30 + i)
        gen_train(model, "nll", b=6, k=10, d=128, regularizer=1e-2 if model == "transe" else 0.0, seed=40 + i)
    for i, model in enumerate(("transe", "distmult", "complex", "simple")):
        gen_eval_loop(model, seed=50 + i)
    gen_normalize()
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
