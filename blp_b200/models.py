"""Host-side mirror of the reference's models.py hot-path surface.

Same names, argument meaning and error behaviour as dfdazac/blp `models.py`
(score functions :222-248, losses :251-266, LinkPrediction :7-70,
InductiveLinkPrediction.forward :78-93, TransductiveLinkPrediction :207-219),
but every computation is a call into libblp_b200.so (sm_100a kernels).  There is
no torch / CPU fallback: CPU tensors raise.

Encoders (BERT, BOW, DKRL) are out of scope and stay the reference's own
PyTorch code; `blp_b200.patch` rebinds the hot path inside the reference's
modules so those classes inherit it.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import lazy, ops


# ---------------------------------------------------------------- score_fn ----
def _score(model, heads, tails, rels):
    """score_fn(heads, tails, rels).  For the eval broadcast of train.py:146-147 -- (1,N,D) against (B,1,D) rows, no
    grad -- with lazy score matrices switched on (blp_b200.patch), the result is a `lazy.LazyScores` handle that
    torch.cat / utils.get_metrics / `pred[mask] = pred.min() - 1` consume without the (B, N) matrix ever existing;
    any other use materialises it with the same exact kernels."""
    handle = lazy.maybe_lazy_score(model, heads, tails, rels)
    return handle if handle is not None else ops.score(model, heads, tails, rels)


def transe_score(heads, tails, rels):
    """models.py:222-223  -||h + r - t||_1 over the last dim."""
    return _score("transe", heads, tails, rels)


def distmult_score(heads, tails, rels):
    """models.py:226-227  sum(h * r * t)."""
    return _score("distmult", heads, tails, rels)


def complex_score(heads, tails, rels):
    """models.py:230-239  Re(<r, h, conj(t)>) on (re | im) halves."""
    return _score("complex", heads, tails, rels)


def simple_score(heads, tails, rels):
    """models.py:242-248  SimplE on (head-role | tail-role) halves, / 2."""
    return _score("simple", heads, tails, rels)


SCORE_FNS = {"transe": transe_score, "distmult": distmult_score, "complex": complex_score, "simple": simple_score}
_SCORE_NAME = {fn: name for name, fn in SCORE_FNS.items()}


# ----------------------------------------------------------------- loss_fn ----
class _PairLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, loss, pos_scores, neg_scores):
        need = pos_scores.requires_grad or neg_scores.requires_grad
        res = ops.pair_loss(loss, pos_scores, neg_scores, want_grad=need)
        if need:
            ctx.save_for_backward(res["grad_pos"], res["grad_neg"])
            ctx.pos_shape = pos_scores.shape
        return res["loss"].reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        gp, gn = ctx.saved_tensors
        g = grad_out.reshape(1).to(torch.float32)
        return None, ops.scaled(gp, g).reshape(ctx.pos_shape), ops.scaled(gn, g)


def margin_loss(pos_scores, neg_scores):
    """models.py:251-254 (margin 1; entries with 1 - pos + neg == 0 keep their gradient)."""
    return _PairLoss.apply("margin", pos_scores, neg_scores)


def nll_loss(pos_scores, neg_scores):
    """models.py:257-258."""
    return _PairLoss.apply("nll", pos_scores, neg_scores)


def l2_regularization(heads, tails, rels):
    """models.py:261-266 (forward only; inside compute_loss the regulariser is fused and differentiable)."""
    return ops.l2_regularization(heads, tails, rels)


LOSS_FNS = {"margin": margin_loss, "nll": nll_loss}
_LOSS_NAME = {fn: name for name, fn in LOSS_FNS.items()}


# ------------------------------------------------------------ compute_loss ----
class _FusedLoss(torch.autograd.Function):
    """models.py:51-70 and its autograd graph as one kernel launch.

    The kernel emits the loss together with d loss / d ent_embs and
    d loss / d rel_emb.weight for an upstream gradient of one; backward scales
    them by grad_output (the loss is linear in it).
    """

    @staticmethod
    def forward(ctx, model, loss, regularizer, ent_embs, rel_weight, rels, neg_idx):
        need = ent_embs.requires_grad or rel_weight.requires_grad
        res = ops.train_loss(model, loss, ent_embs, rel_weight, rels, neg_idx, regularizer=regularizer, want_grad=need)
        if need:
            ctx.save_for_backward(res["grad_all"])
            ctx.ent_shape, ctx.rel_shape = ent_embs.shape, rel_weight.shape
        return res["loss"].reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        (g_all,) = ctx.saved_tensors
        g = ops.scaled(g_all, grad_out.reshape(1).to(torch.float32))        # one launch for both gradients
        n_ent = ctx.ent_shape.numel()
        return None, None, None, g[:n_ent].view(ctx.ent_shape), g[n_ent:].view(ctx.rel_shape), None, None


def fused_compute_loss(model, loss, ent_embs, rel_weight, rels, neg_idx, regularizer=0.0):
    """Functional form of LinkPrediction.compute_loss (models.py:51-70)."""
    return _FusedLoss.apply(model, loss, float(regularizer), ent_embs, rel_weight, rels, neg_idx)


def compute_loss(self, ent_embs, rels, neg_idx):
    """Drop-in body for LinkPrediction.compute_loss (models.py:51-70); `self` may be a reference model."""
    model = getattr(self, "rel_model_name", None) or _SCORE_NAME.get(self.score_fn)
    loss = getattr(self, "loss_fn_name", None) or _LOSS_NAME.get(self.loss_fn)
    if model is None or loss is None:
        raise ValueError("compute_loss: score_fn / loss_fn are not blp_b200 functions; call blp_b200.patch() first")
    return fused_compute_loss(model, loss, ent_embs, self.rel_emb.weight, rels, neg_idx, self.regularizer)


class LinkPrediction(nn.Module):
    """models.py:7-70: relation lookup table + score_fn / loss_fn selected by name."""

    def __init__(self, dim, rel_model, loss_fn, num_relations, regularizer):
        super().__init__()
        self.dim = dim
        self.normalize_embs = False
        self.regularizer = regularizer

        if rel_model not in SCORE_FNS:
            raise ValueError(f'Unknown relational model {rel_model}.')       # models.py:26
        self.score_fn = SCORE_FNS[rel_model]
        self.rel_model_name = rel_model
        if rel_model == 'transe':
            self.normalize_embs = True                                      # models.py:18

        self.rel_emb = nn.Embedding(num_relations, self.dim)
        nn.init.xavier_uniform_(self.rel_emb.weight.data)                   # models.py:28-29

        if loss_fn not in LOSS_FNS:
            raise ValueError(f'Unkown loss function {loss_fn}')             # models.py:36 (sic)
        self.loss_fn = LOSS_FNS[loss_fn]
        self.loss_fn_name = loss_fn

    def encode(self, *args, **kwargs):                                      # models.py:38-43
        ent_emb = self._encode_entity(*args, **kwargs)
        if self.normalize_embs:
            ent_emb = F.normalize(ent_emb, dim=-1)
        return ent_emb

    def _encode_entity(self, *args, **kwargs):
        raise NotImplementedError

    def forward(self, *args, **kwargs):
        raise NotImplementedError

    compute_loss = compute_loss


class InductiveLinkPrediction(LinkPrediction):
    """models.py:73-93: description encoders plug in `_encode_entity`; forward contract unchanged."""

    def _encode_entity(self, text_tok, text_mask):
        raise NotImplementedError

    def forward(self, text_tok, text_mask, rels=None, neg_idx=None):
        batch_size, _, num_text_tokens = text_tok.shape
        ent_embs = self.encode(text_tok.view(-1, num_text_tokens), text_mask.view(-1, num_text_tokens))
        if rels is None and neg_idx is None:
            return ent_embs                                                 # entity embeddings only
        return self.compute_loss(ent_embs.view(batch_size, 2, -1), rels, neg_idx)


class TransductiveLinkPrediction(LinkPrediction):
    """models.py:207-219: entity lookup table."""

    def __init__(self, dim, rel_model, loss_fn, num_entities, num_relations, regularizer):
        super().__init__(dim, rel_model, loss_fn, num_relations, regularizer)
        self.ent_emb = nn.Embedding(num_entities, dim)
        nn.init.xavier_uniform_(self.ent_emb.weight.data)

    def _encode_entity(self, entities):
        return self.ent_emb(entities)

    def forward(self, pos_pairs, rels, neg_idx):
        return self.compute_loss(self.encode(pos_pairs), rels, neg_idx)
