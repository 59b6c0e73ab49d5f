/*
 * blp_oracle.c -- CPU restatement of the BLP scoring / loss / ranking hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package (blp_b200/) may
 * import, link or execute this file.  It is used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
 * legs as the *checker* for the CUDA path.
 *
 * Parity status: PINNED.  The reference (dfdazac/blp) ships no golden vectors,
 * so the pin is the reference itself: oracle/gen_golden.py imports the
 * unmodified /root/reference/{models,utils,train}.py in the build container,
 * runs them on seeded inputs and (a) asserts this restatement is bit-equal on
 * scores / integer ranks, (b) commits the outputs under tests/golden/.
 *
 * Every function cites the reference lines it follows (paths relative to the
 * reference root).  The arithmetic *order* is part of the contract: integer
 * ranks are only reproducible if each fp32 score carries the same roundings
 * as the reference's CPU path (torch ATen, no FMA contraction):
 *
 *   - torch.norm(x, p=1, dim=-1): strictly sequential fp32 sum, j = 0..D-1.
 *   - torch.sum(x, dim=-1) over a contiguous last dim of length L
 *     (ATen vectorized_inner_sum / row_sum / multi_row_sum): 8 vector lanes x
 *     4 interleaved accumulators, cascade levels every 16 steps, remainder
 *     vectors into accumulator 0, accumulators combined 0+1+2+3, scalar tail
 *     first, then lanes 0..7 sequentially.  Verified bit-equal against torch
 *     2.11 for every L in [8, 200) and {256,...,4096} (see gen_golden.py).
 *
 * Build: see oracle/Makefile  (-O2 -ffp-contract=off, OpenMP).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define BLP_TRANSE 0
#define BLP_DISTMULT 1
#define BLP_COMPLEX 2
#define BLP_SIMPLE 3
#define BLP_MARGIN 0
#define BLP_NLL 1

#define ORACLE_MAX_D 8192

int blp_oracle_version(void) { return 1; }

int blp_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void blp_oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---- ATen CPU float sum over a contiguous row of length L ----------------
 * Restates aten/src/ATen/native/cpu/SumKernel.cpp (vectorized_inner_sum ->
 * row_sum -> multi_row_sum) as reached from torch.sum(..., dim=-1) at
 * models.py:227, :235-239, :247.  Vec width 8 floats, ilp_factor 4,
 * num_levels 4, level_power = max(4, CeilLog2(size)/4). */
static int ceil_log2_i64(int64_t x) {
    if (x <= 2) return 1;
    int n = 0;
    int64_t v = x - 1;
    while (v > 0) { v >>= 1; ++n; }
    return n;
}

static float aten_row_sum(const float *p, int L) {
    const int vec_size = L / 8;
    const int size_ilp = vec_size / 4;
    float acc[4][4][8]; /* [level][row][lane] */
    memset(acc, 0, sizeof(acc));

    int level_power = ceil_log2_i64(size_ilp) / 4;
    if (level_power < 4) level_power = 4;
    const int64_t level_step = (int64_t)1 << level_power;
    const int64_t level_mask = level_step - 1;

    int64_t i = 0;
    while (i + level_step <= size_ilp) {
        for (int64_t j = 0; j < level_step; ++j, ++i) {
            const float *base = p + i * 32;
            for (int k = 0; k < 4; ++k)
                for (int l = 0; l < 8; ++l)
                    acc[0][k][l] = acc[0][k][l] + base[k * 8 + l];
        }
        for (int j = 1; j < 4; ++j) {
            for (int k = 0; k < 4; ++k)
                for (int l = 0; l < 8; ++l) {
                    acc[j][k][l] = acc[j][k][l] + acc[j - 1][k][l];
                    acc[j - 1][k][l] = 0.0f;
                }
            const int64_t mask = level_mask << (j * level_power);
            if ((i & mask) != 0) break;
        }
    }
    for (; i < size_ilp; ++i) {
        const float *base = p + i * 32;
        for (int k = 0; k < 4; ++k)
            for (int l = 0; l < 8; ++l)
                acc[0][k][l] = acc[0][k][l] + base[k * 8 + l];
    }
    for (int j = 1; j < 4; ++j)
        for (int k = 0; k < 4; ++k)
            for (int l = 0; l < 8; ++l)
                acc[0][k][l] = acc[0][k][l] + acc[j][k][l];

    /* row_sum: leftover whole vectors go into partial 0, then 0 += 1,2,3 */
    for (int v = size_ilp * 4; v < vec_size; ++v)
        for (int l = 0; l < 8; ++l)
            acc[0][0][l] = acc[0][0][l] + p[v * 8 + l];
    for (int k = 1; k < 4; ++k)
        for (int l = 0; l < 8; ++l)
            acc[0][0][l] = acc[0][0][l] + acc[0][k][l];

    /* vectorized_inner_sum: scalar tail first, then the 8 lanes in order */
    float fin = 0.0f;
    for (int k = vec_size * 8; k < L; ++k) fin = fin + p[k];
    for (int l = 0; l < 8; ++l) fin = fin + acc[0][0][l];
    return fin;
}

float blp_oracle_aten_sum(const float *p, int L) { return aten_row_sum(p, L); }

/* ---- score functions: models.py:222-248 ------------------------------- */

/* models.py:222-223  -norm(heads + rels - tails, p=1): x = fl(fl(h+r)-t),
 * sequential fp32 accumulation of |x|. */
static float transe_score(const float *h, const float *t, const float *r, int d) {
    float s = 0.0f;
    for (int j = 0; j < d; ++j) {
        float x = h[j] + r[j];
        x = x - t[j];
        s = s + fabsf(x);
    }
    return -s;
}

/* models.py:226-227  sum(heads * rels * tails): p = fl(fl(h*r)*t). */
static float distmult_score(const float *h, const float *t, const float *r, int d, float *buf) {
    for (int j = 0; j < d; ++j) {
        float p = h[j] * r[j];
        buf[j] = p * t[j];
    }
    return aten_row_sum(buf, d);
}

/* models.py:230-239  halves via chunk(2,-1); four triple products combined
 * left to right: ((rr*hr*tr + rr*hi*ti) + ri*hr*ti) - ri*hi*tr. */
static float complex_score(const float *h, const float *t, const float *r, int d, float *buf) {
    const int L = d / 2;
    for (int j = 0; j < L; ++j) {
        const float hr = h[j], hi = h[L + j];
        const float tr = t[j], ti = t[L + j];
        const float rr = r[j], ri = r[L + j];
        float a = rr * hr; a = a * tr;
        float b = rr * hi; b = b * ti;
        float c = ri * hr; c = c * ti;
        float e = ri * hi; e = e * tr;
        float p = a + b;
        p = p + c;
        p = p - e;
        buf[j] = p;
    }
    return aten_row_sum(buf, L);
}

/* models.py:242-248  sum(hh*ra*tt + th*rb*ht) / 2. */
static float simple_score(const float *h, const float *t, const float *r, int d, float *buf) {
    const int L = d / 2;
    for (int j = 0; j < L; ++j) {
        const float hh = h[j], ht = h[L + j];
        const float th = t[j], tt = t[L + j];
        const float ra = r[j], rb = r[L + j];
        float a = hh * ra; a = a * tt;
        float b = th * rb; b = b * ht;
        buf[j] = a + b;
    }
    return aten_row_sum(buf, L) / 2.0f;
}

static float score_one(int model, const float *h, const float *t, const float *r, int d, float *buf) {
    switch (model) {
    case BLP_TRANSE: return transe_score(h, t, r, d);
    case BLP_DISTMULT: return distmult_score(h, t, r, d, buf);
    case BLP_COMPLEX: return complex_score(h, t, r, d, buf);
    default: return simple_score(h, t, r, d, buf);
    }
}

static int check_model(int model, int d) {
    if (model < 0 || model > 3) return -1;
    if (d <= 0 || d > ORACLE_MAX_D) return -2;
    if ((model == BLP_COMPLEX || model == BLP_SIMPLE) && (d % 2)) return -3;
    return 0;
}

float blp_oracle_score_one(int model, const float *h, const float *t, const float *r, int d) {
    float buf[ORACLE_MAX_D];
    if (check_model(model, d)) return NAN;
    return score_one(model, h, t, r, d, buf);
}

/* Broadcast scoring over an (A, C) grid of rows, each operand addressed as
 * base + a*sA + c*sC (element strides, 0 = broadcast).  Covers every call
 * shape of score_fn in the reference: train.py:146-147 (eval) and
 * models.py:57,67 (train). */
int blp_oracle_score_bcast(int model,
                           const float *h, int64_t hsA, int64_t hsC,
                           const float *t, int64_t tsA, int64_t tsC,
                           const float *r, int64_t rsA, int64_t rsC,
                           int64_t A, int64_t C, int d, float *out) {
    int rc = check_model(model, d);
    if (rc) return rc;
#pragma omp parallel
    {
        float *buf = (float *)malloc(sizeof(float) * (size_t)d);
#pragma omp for collapse(2) schedule(static)
        for (int64_t a = 0; a < A; ++a)
            for (int64_t c = 0; c < C; ++c)
                out[a * C + c] = score_one(model, h + a * hsA + c * hsC, t + a * tsA + c * tsC,
                                           r + a * rsA + c * rsC, d, buf);
        free(buf);
    }
    return 0;
}

/* ---- utils.py:86-111 get_metrics: integer part -------------------------
 * gt = #{j: s_j > s_true}; ge = #{j: s_j >= s_true}  (best_rank = gt + 1,
 * worst_rank = ge). */
int blp_oracle_rank_counts(const float *pred, int64_t q, int64_t n, const int64_t *true_idx,
                           int64_t *gt, int64_t *ge) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < q; ++i) {
        const float *row = pred + i * n;
        const float st = row[true_idx[i]];
        int64_t a = 0, b = 0;
        for (int64_t j = 0; j < n; ++j) {
            a += row[j] > st;
            b += row[j] >= st;
        }
        gt[i] = a;
        ge[i] = b;
    }
    return 0;
}

/* utils.py:106-109: average_rank = (best + worst).float() * 0.5;
 * reciprocals = 1/average_rank; hits = average_rank <= k. */
int blp_oracle_metrics_from_counts(const int64_t *gt, const int64_t *ge, int64_t q,
                                   const int64_t *k_values, int nk, float *recip, uint8_t *hits) {
    for (int64_t i = 0; i < q; ++i) {
        const float avg = (float)(gt[i] + 1 + ge[i]) * 0.5f;
        recip[i] = 1.0f / avg;
        for (int j = 0; j < nk; ++j) hits[i * nk + j] = avg <= (float)k_values[j];
    }
    return 0;
}

/* ---- train.py:141-171: full-entity scoring + raw/filtered rank counts --
 * Query i < B predicts the head (candidate row plays `heads`, train.py:146);
 * query B+i predicts the tail (candidate row plays `tails`, train.py:147);
 * cat order heads-then-tails (train.py:149-150).  Filtered ranks
 * (train.py:159-167): filtered candidates are pushed below every score, so
 * they leave both counts; the true entity is never filtered (utils.py:71,78).
 * filt_indptr[2B+1]/filt_idx: CSR of filtered candidate rows per query
 * (unique per query), or NULL.  scores_out (2B*N) optional. */
int blp_oracle_eval_rank(int model, const float *ent, int64_t n, int d,
                         const float *h_rows, const float *t_rows, const float *r_rows,
                         const int64_t *head_idx, const int64_t *tail_idx, int64_t b,
                         const int64_t *filt_indptr, const int64_t *filt_idx,
                         int64_t *gt, int64_t *ge, int64_t *gt_f, int64_t *ge_f,
                         float *true_score, float *scores_out) {
    int rc = check_model(model, d);
    if (rc) return rc;
#pragma omp parallel
    {
        float *buf = (float *)malloc(sizeof(float) * (size_t)d);
#pragma omp for schedule(dynamic, 1)
        for (int64_t q = 0; q < 2 * b; ++q) {
            const int head_pred = q < b;
            const int64_t i = head_pred ? q : q - b;
            const float *hq = h_rows + i * d, *tq = t_rows + i * d, *rq = r_rows + i * d;
            /* pred.gather(true_idx) (utils.py:103).  With head_idx / tail_idx == NULL the true row is taken
             * from the gathered query rows themselves (h_rows[i] IS ent[head_idx[i]], train.py:141-142), which
             * is what a candidate-sharded caller must do when the true row lives in another shard. */
            const float *e_true = head_pred ? (head_idx ? ent + head_idx[i] * d : hq)
                                            : (tail_idx ? ent + tail_idx[i] * d : tq);
            const float st = head_pred ? score_one(model, e_true, tq, rq, d, buf)
                                       : score_one(model, hq, e_true, rq, d, buf);
            int64_t a = 0, c = 0;
            for (int64_t j = 0; j < n; ++j) {
                const float *e = ent + j * d;
                const float s = head_pred ? score_one(model, e, tq, rq, d, buf)
                                          : score_one(model, hq, e, rq, d, buf);
                if (scores_out) scores_out[q * n + j] = s;
                a += s > st;
                c += s >= st;
            }
            gt[q] = a;
            ge[q] = c;
            if (true_score) true_score[q] = st;
            if (gt_f && ge_f) {
                int64_t fa = a, fc = c;
                if (filt_indptr) {
                    for (int64_t p = filt_indptr[q]; p < filt_indptr[q + 1]; ++p) {
                        const float *e = ent + filt_idx[p] * d;
                        const float s = head_pred ? score_one(model, e, tq, rq, d, buf)
                                                  : score_one(model, hq, e, rq, d, buf);
                        fa -= s > st;
                        fc -= s >= st;
                    }
                }
                gt_f[q] = fa;
                ge_f[q] = fc;
            }
        }
        free(buf);
    }
    return 0;
}

/* ---- models.py:51-70 compute_loss + models.py:251-266 losses -----------
 * ent_embs (B,2,D) contiguous; rel_rows (B,D) = rel_emb(rels) (models.py:55);
 * neg_idx (B,K,2) int64 addressed with element strides (s0,s1,s2) because the
 * reference sampler returns a transposed view (data.py:78-79); values index
 * ent_embs.view(2B, D) (models.py:65).
 * Scores are bit-exact restatements; the scalar loss is accumulated in double
 * (the reference's mean() is a threaded full reduction whose order is not
 * stable, so the loss is a tolerance quantity).  Gradients are analytic, in
 * double, following autograd's conventions for the reference ops:
 * margin: d/dm of the masked `loss[loss<0]=0` is 1 where m >= 0 (models.py:253);
 * transe: sign(0) = 0; softplus threshold 20 (models.py:258).
 * grad_ent (B,2,D), grad_rel (B,D) may be NULL (forward only). */
static double softplus_d(double x) { return x > 20.0 ? x : log1p(exp(x)); }
static double softplus_grad_d(double x) {
    if (x > 20.0) return 1.0;
    const double z = exp(x);
    return z / (z + 1.0);
}

static void score_grad(int model, const float *h, const float *t, const float *r, int d, double w,
                       double *gh, double *gt_, double *gr) {
    if (model == BLP_TRANSE) {
        for (int j = 0; j < d; ++j) {
            float x = h[j] + r[j];
            x = x - t[j];
            const double sg = (x > 0.0f) - (x < 0.0f);
            gh[j] -= w * sg;
            gr[j] -= w * sg;
            gt_[j] += w * sg;
        }
    } else if (model == BLP_DISTMULT) {
        for (int j = 0; j < d; ++j) {
            gh[j] += w * (double)r[j] * t[j];
            gt_[j] += w * (double)h[j] * r[j];
            gr[j] += w * (double)h[j] * t[j];
        }
    } else if (model == BLP_COMPLEX) {
        const int L = d / 2;
        for (int j = 0; j < L; ++j) {
            const double hr = h[j], hi = h[L + j], tr = t[j], ti = t[L + j], rr = r[j], ri = r[L + j];
            gh[j] += w * (rr * tr + ri * ti);
            gh[L + j] += w * (rr * ti - ri * tr);
            gt_[j] += w * (rr * hr - ri * hi);
            gt_[L + j] += w * (rr * hi + ri * hr);
            gr[j] += w * (hr * tr + hi * ti);
            gr[L + j] += w * (hr * ti - hi * tr);
        }
    } else {
        const int L = d / 2;
        for (int j = 0; j < L; ++j) {
            const double hh = h[j], ht = h[L + j], th = t[j], tt = t[L + j], ra = r[j], rb = r[L + j];
            gh[j] += 0.5 * w * ra * tt;
            gh[L + j] += 0.5 * w * th * rb;
            gt_[j] += 0.5 * w * rb * ht;
            gt_[L + j] += 0.5 * w * hh * ra;
            gr[j] += 0.5 * w * hh * tt;
            gr[L + j] += 0.5 * w * th * ht;
        }
    }
}

int blp_oracle_train_loss(int model, int loss, const float *ent_embs, const float *rel_rows,
                          const int64_t *neg_idx, int64_t s0, int64_t s1, int64_t s2,
                          int64_t b, int64_t k, int d, float regularizer, float grad_out,
                          float *loss_out, float *pos_scores, float *neg_scores,
                          float *grad_ent, float *grad_rel) {
    int rc = check_model(model, d);
    if (rc) return rc;
    if (loss != BLP_MARGIN && loss != BLP_NLL) return -4;
    float *buf = (float *)malloc(sizeof(float) * (size_t)d);
    float *pos = (float *)malloc(sizeof(float) * (size_t)b);
    float *neg = (float *)malloc(sizeof(float) * (size_t)(b * k));
    /* models.py:57 positive scores; models.py:65-67 negative scores */
    for (int64_t i = 0; i < b; ++i)
        pos[i] = score_one(model, ent_embs + (2 * i) * d, ent_embs + (2 * i + 1) * d, rel_rows + i * d, d, buf);
    for (int64_t i = 0; i < b; ++i)
        for (int64_t j = 0; j < k; ++j) {
            const int64_t i0 = neg_idx[i * s0 + j * s1], i1 = neg_idx[i * s0 + j * s1 + s2];
            if (i0 < 0 || i0 >= 2 * b || i1 < 0 || i1 >= 2 * b) { free(buf); free(pos); free(neg); return -5; }
            neg[i * k + j] = score_one(model, ent_embs + i0 * d, ent_embs + i1 * d, rel_rows + i * d, d, buf);
        }
    double total = 0.0;
    if (loss == BLP_MARGIN) {
        /* models.py:251-254: m = fl(fl(1 - pos) + neg); zero where m < 0; mean */
        for (int64_t i = 0; i < b; ++i) {
            const float one_minus = 1.0f - pos[i];
            for (int64_t j = 0; j < k; ++j) {
                const float m = one_minus + neg[i * k + j];
                if (!(m < 0.0f)) total += (double)m;
            }
        }
        total /= (double)(b * k);
    } else {
        /* models.py:257-258 */
        double sp = 0.0, sn = 0.0;
        for (int64_t i = 0; i < b; ++i) sp += softplus_d(-(double)pos[i]);
        for (int64_t i = 0; i < b * k; ++i) sn += softplus_d((double)neg[i]);
        total = (sp / (double)b + sn / (double)(b * k)) / 2.0;
    }
    /* models.py:59-62, 261-266: regularizer * (mean h^2 + mean t^2 + mean r^2)/3 on the positives */
    if (regularizer > 0.0f) {
        double sh = 0.0, st = 0.0, sr = 0.0;
        for (int64_t i = 0; i < b; ++i)
            for (int j = 0; j < d; ++j) {
                const double hv = ent_embs[(2 * i) * d + j], tv = ent_embs[(2 * i + 1) * d + j], rv = rel_rows[i * d + j];
                sh += hv * hv; st += tv * tv; sr += rv * rv;
            }
        const double nrm = (double)b * d;
        total += (double)regularizer * ((sh / nrm + st / nrm + sr / nrm) / 3.0);
    }
    if (loss_out) *loss_out = (float)total;
    if (pos_scores) memcpy(pos_scores, pos, sizeof(float) * (size_t)b);
    if (neg_scores) memcpy(neg_scores, neg, sizeof(float) * (size_t)(b * k));

    if (grad_ent && grad_rel) {
        const size_t ne = (size_t)(2 * b) * d, nr = (size_t)b * d;
        double *ge_ = (double *)calloc(ne, sizeof(double));
        double *gr_ = (double *)calloc(nr, sizeof(double));
        const double g = (double)grad_out;
        for (int64_t i = 0; i < b; ++i) {
            double wpos = 0.0;
            const float one_minus = 1.0f - pos[i];
            for (int64_t j = 0; j < k; ++j) {
                double w;
                if (loss == BLP_MARGIN) {
                    const float m = one_minus + neg[i * k + j];
                    w = (m < 0.0f) ? 0.0 : g / (double)(b * k);
                    wpos -= w;
                } else {
                    w = g * softplus_grad_d((double)neg[i * k + j]) / (2.0 * (double)(b * k));
                }
                if (w != 0.0) {
                    const int64_t i0 = neg_idx[i * s0 + j * s1], i1 = neg_idx[i * s0 + j * s1 + s2];
                    score_grad(model, ent_embs + i0 * d, ent_embs + i1 * d, rel_rows + i * d, d, w,
                               ge_ + i0 * d, ge_ + i1 * d, gr_ + i * d);
                }
            }
            if (loss == BLP_NLL) wpos = -g * softplus_grad_d(-(double)pos[i]) / (2.0 * (double)b);
            if (wpos != 0.0)
                score_grad(model, ent_embs + (2 * i) * d, ent_embs + (2 * i + 1) * d, rel_rows + i * d, d, wpos,
                           ge_ + (2 * i) * d, ge_ + (2 * i + 1) * d, gr_ + i * d);
        }
        if (regularizer > 0.0f) {
            const double c = g * (double)regularizer * 2.0 / (3.0 * (double)b * d);
            for (size_t j = 0; j < ne; ++j) ge_[j] += c * (double)ent_embs[j];
            for (size_t j = 0; j < nr; ++j) gr_[j] += c * (double)rel_rows[j];
        }
        for (size_t j = 0; j < ne; ++j) grad_ent[j] = (float)ge_[j];
        for (size_t j = 0; j < nr; ++j) grad_rel[j] = (float)gr_[j];
        free(ge_);
        free(gr_);
    }
    free(buf); free(pos); free(neg);
    return 0;
}
