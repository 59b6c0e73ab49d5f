"""ATen-op restatement of the reference's CPU hot path (the CPU *baseline*).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Where blp_oracle.c and
np_oracle.py restate the arithmetic, this module restates the reference's
*execution strategy*: the same stock torch ops in the same order, including
the materialised (B, N, D) broadcasts, so that timing it on a host measures
what the reference's own CPU path costs there (/root/reference itself is not
available on the GPU box).  Used as bench.py's `cpu_baseline` / `--impl
reference` leg and as a third cross-check of the oracle.
"""
import torch
import torch.nn.functional as F


def _score_l1(h, t, r):          # models.py:222-223
    return -torch.norm(h + r - t, dim=-1, p=1)


def _score_trilinear(h, t, r):   # models.py:226-227
    return torch.sum(h * r * t, dim=-1)


def _score_complex(h, t, r):     # models.py:230-239
    h_re, h_im = torch.chunk(h, 2, dim=-1)
    t_re, t_im = torch.chunk(t, 2, dim=-1)
    r_re, r_im = torch.chunk(r, 2, dim=-1)
    return torch.sum(r_re * h_re * t_re + r_re * h_im * t_im + r_im * h_re * t_im - r_im * h_im * t_re, dim=-1)


def _score_simple(h, t, r):      # models.py:242-248
    h_h, h_t = torch.chunk(h, 2, dim=-1)
    t_h, t_t = torch.chunk(t, 2, dim=-1)
    r_a, r_b = torch.chunk(r, 2, dim=-1)
    return torch.sum(h_h * r_a * t_t + t_h * r_b * h_t, dim=-1) / 2


SCORE_FNS = {"transe": _score_l1, "distmult": _score_trilinear, "complex": _score_complex, "simple": _score_simple}


def hinge(pos, neg):             # models.py:251-254
    out = 1 - pos + neg
    out[out < 0] = 0
    return out.mean()


def logistic(pos, neg):          # models.py:257-258
    return (F.softplus(-pos).mean() + F.softplus(neg).mean()) / 2


LOSS_FNS = {"margin": hinge, "nll": logistic}


def rank_metrics(pred, true_idx, k_values):   # utils.py:86-111
    true = pred.gather(dim=1, index=true_idx)
    best = (pred > true).sum(dim=1, keepdim=True) + 1
    worst = (pred >= true).sum(dim=1, keepdim=True)
    avg = (best + worst).float() * 0.5
    return avg.reciprocal(), avg <= k_values


def batch_loss(model, loss, ent_embs, rel_rows, neg_idx, regularizer=0.0):   # models.py:51-70
    b = ent_embs.shape[0]
    rels = rel_rows.unsqueeze(1)
    heads, tails = torch.chunk(ent_embs, 2, dim=1)
    fn = SCORE_FNS[model]
    pos = fn(heads, tails, rels)
    reg = 0
    if regularizer > 0:
        reg = regularizer * sum(torch.mean(x ** 2) for x in (heads, tails, rels)) / 3.0
    neg_embs = ent_embs.view(b * 2, -1)[neg_idx]
    nh, nt = torch.chunk(neg_embs, 2, dim=2)
    neg = fn(nh.squeeze(), nt.squeeze(), rels)
    return LOSS_FNS[loss](pos, neg) + reg


@torch.no_grad()
def eval_batch(model, ent_emb, heads, tails, rel_rows, k_values, filter_mask=None):   # train.py:141-171
    """ent_emb (1,N,D); heads/tails (B,1) long; rel_rows (B,1,D); -> metrics like the reference loop."""
    fn = SCORE_FNS[model]
    table = ent_emb.squeeze(0)
    head_embs, tail_embs = table[heads], table[tails]
    pred = torch.cat((fn(ent_emb, tail_embs, rel_rows), fn(head_embs, ent_emb, rel_rows)))
    true = torch.cat((heads, tails))
    recip, hits = rank_metrics(pred, true, k_values)
    out = {"pred": pred, "recip": recip, "hits": hits}
    if filter_mask is not None:
        pred[filter_mask] = pred.min() - 1.0
        out["recip_f"], out["hits_f"] = rank_metrics(pred, true, k_values)
    return out


# ---- CPU restatements of the rows either side of the hot path (timing baselines for SURVEY section 8f) ----
def sample_negative_indices(batch_size, num_negatives, repeats=1):   # data.py:35-81
    """In-batch corruption indices the way the reference draws them: a (B, 2B) weight matrix that is zero on the
    own pair, torch.multinomial with replacement, a coin per negative for the corrupted side; returns the
    transposed (B * repeats, K, 2) view."""
    n = 2 * batch_size
    own = torch.arange(n).view(batch_size, 2)
    weights = torch.ones(batch_size, n)
    weights.scatter_(1, own, torch.zeros(batch_size, 2))
    draws = weights.multinomial(num_negatives * repeats, replacement=True).t().reshape(-1)
    total = batch_size * num_negatives * repeats
    side = torch.randint(0, 2, [total])
    out = own.repeat((num_negatives * repeats, 1))
    out[torch.arange(total), side] = draws
    return out.view(-1, batch_size * repeats, 2).transpose(0, 1)


def triple_filter_masks(triples, out_edges, in_edges, num_ents, ent2idx):   # utils.py:46-83
    """Dense (B, N) bool masks of known tails / heads per test triple, built with the reference's per-edge Python loops.
    out_edges[h] / in_edges[t] list the (head, tail, rel) edges of the filtering graph (what graph.out_edges /
    graph.in_edges iterate)."""
    b = triples.shape[0]
    heads_mask = torch.zeros((b, num_ents), dtype=torch.bool)
    tails_mask = torch.zeros((b, num_ents), dtype=torch.bool)
    for i, (head, tail, rel) in enumerate(triples.tolist()):
        for (_, t, r) in out_edges.get(head, ()):
            if r == rel and t != tail and ent2idx[t] != -1:
                tails_mask[i, ent2idx[t]] = True
        for (h, _, r) in in_edges.get(tail, ()):
            if r == rel and h != head and ent2idx[h] != -1:
                heads_mask[i, ent2idx[h]] = True
    return heads_mask, tails_mask


def mrr_by_new_position(triples, recip, new_entities):   # utils.py:114-147
    sums, counts = torch.zeros(3), torch.zeros(3)
    n = triples.shape[0]
    for i, (h, t, _) in enumerate(triples):
        head, tail = h.item(), t.item()
        v = (recip[i] + recip[i + n]).item() / 2.0
        slot = 0 if (head in new_entities and tail in new_entities) else 1 if head in new_entities else 2 if tail in new_entities else None
        if slot is not None:
            sums[slot] += v
            counts[slot] += 1.0
    return sums, counts


def normalize_rows(x):   # models.py:40-41
    return F.normalize(x, dim=-1)
