/*
 * blp_b200.h -- C ABI of the B200-native BLP scoring / loss / ranking path.
 *
 * The reference (dfdazac/blp) has no FFI: its seam is Python attribute binding
 * (models.py:16-24 `self.score_fn = transe_score`, models.py:31-34
 * `self.loss_fn = margin_loss`, utils.get_metrics called at train.py:153,167).
 * This header is therefore the boundary a maintainer binds with ctypes (see
 * INTEGRATION.md); each entry point names the reference lines it replaces.
 *
 * Conventions
 *   - every pointer is a *borrowed device pointer* (cudaMalloc'd memory on the
 *     current device) unless marked "host"; the library allocates nothing;
 *   - `stream` is a cudaStream_t passed as void*; all work is asynchronous on
 *     it, no call synchronises or copies to/from the host;
 *   - return 0 on success, a negative BLP_E* code otherwise; the message is in
 *     the thread-local blp_last_error();
 *   - re-entrant: no global mutable state (DataParallel calls forward() from
 *     one Python thread per device, train.py:330);
 *   - fp32 only; rows are contiguous with `d` floats; D even for complex/simple
 *     (torch.chunk(2, -1), models.py:231-233, 243-245).
 */
#ifndef BLP_B200_H
#define BLP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BLP_B200_VERSION 100

/* rel_model, models.py:16-26 */
#define BLP_MODEL_TRANSE 0
#define BLP_MODEL_DISTMULT 1
#define BLP_MODEL_COMPLEX 2
#define BLP_MODEL_SIMPLE 3
/* loss_fn, models.py:31-36 */
#define BLP_LOSS_MARGIN 0
#define BLP_LOSS_NLL 1

#define BLP_OK 0
#define BLP_EINVAL (-1)   /* bad model / loss / shape / null pointer */
#define BLP_EDIM (-2)     /* d not supported (odd d for complex/simple, d <= 0) */
#define BLP_ECUDA (-3)    /* CUDA runtime error (message has the cudaError string) */
#define BLP_EARCH (-4)    /* device is not sm_100 */

int blp_version(void);
const char *blp_last_error(void);
/* 0 if `device` is a compute-capability 10.x GPU this library was built for. */
int blp_device_check(int device);

/* ---- a1-a4  score_fn(heads, tails, rels)  (models.py:222-248) ------------
 * Materialising form, for direct callers of score_fn.  Operands are viewed as
 * (A, C, d) with element strides (sA, sC) per operand, 0 = broadcast; covers
 * the eval shapes (1,N,D)x(B,1,D)x(B,1,D) (train.py:146-147) and the train
 * shapes (B,K,D)x(B,K,D)x(B,1,D) (models.py:57,67).  out is (A, C) row-major.
 * Every score carries the same fp32 roundings as the reference's CPU path. */
int blp_score_bcast(int model,
                    const float *heads, int64_t hsA, int64_t hsC,
                    const float *tails, int64_t tsA, int64_t tsC,
                    const float *rels, int64_t rsA, int64_t rsC,
                    int64_t A, int64_t C, int d, float *out, void *stream);

/* ---- a11  utils.get_metrics  (utils.py:86-111) ----------------------------
 * Integer part on a materialised (q, n) score matrix with row stride
 * `row_stride` floats: gt[i] = #{j: s_ij > s_i,true}, ge[i] = #{j: s_ij >= ..}
 * (best_rank = gt + 1, worst_rank = ge). */
int blp_rank_counts(const float *pred, int64_t q, int64_t n, int64_t row_stride,
                    const int64_t *true_idx, int32_t *gt, int32_t *ge, void *stream);
/* Float part: avg = float(gt + 1 + ge) * 0.5; recip = 1/avg; hits = avg <= k.
 * k_values is a HOST array of nk ints (nk <= 8); hits is (q, nk) bytes. */
int blp_metrics_from_counts(const int32_t *gt, const int32_t *ge, int64_t q,
                            const int64_t *k_values_host, int nk,
                            float *recip, uint8_t *hits, void *stream);

/* train.py:154-157: sums[0] = sum_i 1/avg_i, sums[1 + j] = #{i: avg_i <= k_j}
 * over q queries, accumulated in fp64 in a fixed order.  sums is a DEVICE array
 * of 1 + nk doubles (overwritten); k_values is a HOST array. */
int blp_metrics_reduce(const int32_t *gt, const int32_t *ge, int64_t q,
                       const int64_t *k_values_host, int nk, double *sums, void *stream);

/* ---- a10 + a11 + a12  fused full-entity scoring and ranking ---------------
 * Replaces train.py:141-171 for one batch (or a whole sweep: b is unbounded):
 * queries 0..b-1 predict the head (every candidate row plays `heads`,
 * train.py:146), queries b..2b-1 predict the tail (train.py:147); no
 * (2b, n) score matrix is written.
 *   ent        [n_local, d]  this rank's shard of the entity table
 *   ent_offset global row id of ent[0] (filter indices are global)
 *   h_rows, t_rows, r_rows [b, d]  gathered true head / true tail / relation
 *              rows (train.py:141-143); the true entity's score is computed
 *              from these rows, so every shard derives identical bits
 *   filt_indptr [2b+1], filt_idx [nnz]  CSR of filtered candidate rows per
 *              query (global ids, unique per query; the true entity is never
 *              listed, utils.py:71,78), or NULL for raw ranks only
 *   gt, ge     [2b] raw counts over this shard (overwritten)
 *   gt_f, ge_f [2b] filtered counts (overwritten; may be NULL iff filt is NULL)
 *   true_score [2b] score of the true triple (overwritten)
 * Counts are integers, so sums over shards are exact for any partition. */
int blp_eval_rank(int model, const float *ent, int64_t n_local, int64_t ent_offset, int d,
                  const float *h_rows, const float *t_rows, const float *r_rows, int64_t b,
                  const int64_t *filt_indptr, const int64_t *filt_idx,
                  int32_t *gt, int32_t *ge, int32_t *gt_f, int32_t *ge_f,
                  float *true_score, void *stream);

/* ---- a10 + a11 + a12 for a whole sweep, query rows gathered in-kernel ------
 * Same computation as blp_eval_rank, but the per-batch gathers of train.py:141-143
 * (ent_emb[heads], ent_emb[tails], rel_emb(rels)) are folded into the kernels:
 *   triples    [t, 3] int64 (head row, tail row, relation id): GLOBAL table rows, i.e.
 *              after ent2idx (train.py:132-135); an id outside the table / relation range
 *              gives that triple a NaN true_score and zero counts
 *   rel_weight [num_rel, d]
 *   h_rows, t_rows  optional pre-gathered [t, d] true rows (entity-sharded sweeps, where a
 *              true row may live on another rank); NULL = gather from `ent`
 *   tail_off   output slot of tail query i is tail_off + i (head query i -> slot i), so a
 *              chunk of a longer sweep writes straight into (2, T) arrays: tail_off >= t
 *   filt_indptr [2t+1]: heads' rows first, then tails' rows (chunk-local), as blp_eval_rank
 * Launches: true scores, sweep, (filter correction). */
int blp_rank_sweep(int model, const float *ent, int64_t n_local, int64_t ent_offset, int d,
                   const float *rel_weight, int64_t num_rel, const int64_t *triples, int64_t t,
                   const float *h_rows, const float *t_rows,
                   const int64_t *filt_indptr, const int64_t *filt_idx, int64_t tail_off,
                   int32_t *gt, int32_t *ge, int32_t *gt_f, int32_t *ge_f, float *true_score, void *stream);

/* The two phases of blp_rank_sweep as separate calls, for sweeps processed in chunks (the reference's eval
 * batches, train.py:128): blp_true_scores writes true_score and zeroes gt / ge for ALL t triples of the sweep in
 * one launch; blp_rank_sweep_counts then counts one chunk per call (it expects true_score / zeroed counters in
 * place and launches only the sweep, plus the CSR filter correction when filters are given). */
int blp_true_scores(int model, const float *ent, int64_t n_local, int64_t ent_offset, int d,
                    const float *rel_weight, int64_t num_rel, const int64_t *triples, int64_t t,
                    const float *h_rows, const float *t_rows, int64_t tail_off, int32_t *gt, int32_t *ge,
                    float *true_score, void *stream);
int blp_rank_sweep_counts(int model, const float *ent, int64_t n_local, int64_t ent_offset, int d,
                          const float *rel_weight, int64_t num_rel, const int64_t *triples, int64_t t,
                          const float *h_rows, const float *t_rows,
                          const int64_t *filt_indptr, const int64_t *filt_idx, int64_t tail_off,
                          int32_t *gt, int32_t *ge, int32_t *gt_f, int32_t *ge_f, const float *true_score, void *stream);

/* ---- the fused step: ONE launch per eval batch (train.py:141-157 for one batch of the reference's loop) ----
 * As blp_rank_sweep without filter lists, plus the metrics: the true scores are computed per query group inside
 * the sweep kernel, the counters accumulate in `workspace`, and the last CTA to finish writes gt / ge and --
 * when sums != NULL -- utils.get_metrics' reciprocal ranks / hits (utils.py:106-109; recip [tail_off + t],
 * hits [(tail_off + t) * nk], either may be NULL) and the train.py:154-157 accumulators sums[1 + nk] (fp64:
 * sum of reciprocal ranks, hit counts).  d == 128 on a non-empty shard; other shapes run the separate kernels.
 *   group_triples  triples per table pass: 0 = chosen from t; 2 / 4 / 8 / 16 / 32 forces it.  The reference's
 *                  eval_batch_size is exactly this quantity (each batch streams the whole table once), so a
 *                  Wikidata5M-style sweep at eval batch 2 is ONE call with t = all test triples, group_triples = 2.
 *   workspace      blp_rank_step_workspace_bytes(tail_off + t) bytes, ZERO on first use; the kernel leaves it
 *                  zeroed.  One per stream.  NULL (or tail_off + t > 16384): memset + separate metrics launch.
 *   multi-GPU      pass sums = NULL, all-reduce gt / ge over the shards, then blp_rank_metrics. */
int64_t blp_rank_step_workspace_bytes(int64_t out_len);
int blp_rank_step(int model, const float *ent, int64_t n_local, int64_t ent_offset, int d,
                  const float *rel_weight, int64_t num_rel, const int64_t *triples, int64_t t,
                  const float *h_rows, const float *t_rows, int64_t tail_off, int group_triples,
                  int32_t *gt, int32_t *ge, float *true_score, const int64_t *k_values_host, int nk,
                  float *recip, uint8_t *hits, double *sums, void *workspace, void *stream);
/* blp_rank_step with everything but the batch bound once: `plan` is a host object owned by the caller
 * (blp_plan_destroy frees it); blp_plan_run(plan, triples, h_rows, t_rows, stream) is one step.  The reference
 * calls the step every eval_batch_size = 64 triples (train.py:128), where argument marshalling is a visible
 * share of a ~20 us step. */
int blp_plan_create(void **plan_out, int model, const float *ent, int64_t n_local, int64_t ent_offset, int d,
                    const float *rel_weight, int64_t num_rel, int64_t t, int64_t tail_off, int group_triples,
                    int32_t *gt, int32_t *ge, float *true_score, const int64_t *k_values_host, int nk,
                    float *recip, uint8_t *hits, double *sums, void *workspace);
/* blp_plan_set_overlap(plan, 1): consecutive blp_plan_run calls on one stream overlap -- the step is launched with
 * programmatic stream serialization, so the CTAs of batch n + 1 take the SMs as the CTAs of batch n exit, run their
 * query prologue and scoring, and only wait for batch n to complete before they touch the workspace / outputs
 * (launch latency, the prologue and the work-item quantisation tail of batch n are hidden: 64-triple FB15k-237
 * batches 36.5 -> ~27 us per call).  CONTRACT: everything the call reads (table, relation table, triples, h_rows /
 * t_rows) must be complete when the call is enqueued -- produced by host copies or by work that finished earlier,
 * NOT by a kernel launched immediately before it on the same stream (slices of a resident triple tensor are fine).
 * Outputs of batch n are overwritten by batch n + 1 as before.  Off by default. */
int blp_plan_set_overlap(void *plan, int on);
int blp_plan_run(void *plan, const int64_t *triples, const float *h_rows, const float *t_rows, void *stream);
void blp_plan_destroy(void *plan);

/* ---- get_metrics(score_fn(...), true_idx, k) without the score matrix (train.py:146-153 left untouched) ----
 * Ranks unrelated queries against the table: n_hq head-prediction queries -- candidate row e scored as
 * score_fn(e, hq_tails[i], hq_rels[i]) (train.py:146), true candidate row hq_true[i] (global id) -- followed by
 * n_tq tail-prediction queries score_fn(tq_heads[i], e, tq_rels[i]) (train.py:147) with true candidate tq_true[i].
 * Query rows are dense [n, d]; outputs (gt, ge, true_score, recip, hits) hold the head queries first, i.e. the
 * row order of torch.cat((heads_predictions, tails_predictions)) (train.py:149).  One launch when n_hq == n_tq,
 * d == 128 and a workspace (blp_rank_step_workspace_bytes(n_hq + n_tq)) is given. */
int blp_rank_queries(int model, const float *ent, int64_t n_local, int64_t ent_offset, int d,
                     const float *hq_tails, const float *hq_rels, const int64_t *hq_true, int64_t n_hq,
                     const float *tq_heads, const float *tq_rels, const int64_t *tq_true, int64_t n_tq,
                     int32_t *gt, int32_t *ge, float *true_score, const int64_t *k_values_host, int nk,
                     float *recip, uint8_t *hits, double *sums, void *workspace, void *stream);

/* ---- tensor-core ("fast") mode of the sweep: distmult / complex / simple, d = 128 ----
 * The bilinear scores are linear in the candidate row once the query side is folded
 * (models.py:226-248 with the candidate factored out), so the sweep is a (2t x 128) x (128 x n)
 * contraction: tcgen05.mma kind::f16 on split-FP16 operands (x scaled by a power of two, then hi = fp16(x),
 * lo = fp16(x - hi); hi*hi + lo*hi + hi*lo accumulated in fp32 in TMEM -- 22 significant bits, like
 * 3xTF32 at twice the MMA rate and half the operand bytes) and the rank-count epilogue reading TMEM.
 * NOT bit-exact (folding and the summation order differ from models.py:227): scores agree to
 * ~1e-6 * sum|terms| and ranks differ only for candidates inside that band around the true score.
 * The true score itself is the exact one, and the true entity always counts as a tie (utils.py:104-105).
 *   table_ws   blp_fast_table_bytes(n_local) bytes, filled by blp_fast_prepare_table (once per
 *              table: max-|e| scale + split); query_ws  blp_fast_query_bytes(t) bytes of scratch
 *   scores_out optional (2t, ld_scores) matrix receiving the fast scores (row q: head queries
 *              then tail queries; column: local candidate) -- verification aid, NULL normally
 * Other arguments as blp_rank_sweep. */
int64_t blp_fast_table_bytes(int64_t n_local);
int64_t blp_fast_query_bytes(int64_t t);
int blp_fast_prepare_table(const float *ent, int64_t n_local, int d, void *table_ws, void *stream);
int blp_rank_sweep_fast(int model, const float *ent, int64_t n_local, int64_t ent_offset, int d,
                        const float *rel_weight, int64_t num_rel, const int64_t *triples, int64_t t,
                        const float *h_rows, const float *t_rows,
                        const int64_t *filt_indptr, const int64_t *filt_idx, int64_t tail_off,
                        int32_t *gt, int32_t *ge, int32_t *gt_f, int32_t *ge_f, float *true_score,
                        const void *table_ws, void *query_ws, float *scores_out, int64_t ld_scores,
                        void *stream);

/* ---- exact ranks on the tensor path: filter + refine (distmult / complex / simple, d = 128) ----
 * blp_rank_sweep_fast with an a-priori error band around the true score.  The fold kernel derives, per query,
 *   band = kappa * || |folded coefficients| ||_2 * max_e ||e||_2      (kappa = 2e-5, scaled like the accumulator)
 * which bounds |fast score - reference fp32 score| for every candidate (Cauchy-Schwarz over sum|terms|; the
 * reference's own rounding, the fold, the split and the tensor-core accumulation are each <= ~8e-6 * sum|terms|,
 * DESIGN.md 4.2b).  The sweep kernel counts only candidates that beat the true score by more than the band; the
 * few candidates inside the band (a handful per query) are appended to a worklist and re-scored by a refine
 * kernel with the reference's fp32 operations in the reference's summation order (models.py:226-248), the same
 * code that produces the true-triple scores.  gt / ge (and the filtered counters) are then bit-identical to
 * blp_rank_sweep (exact mode) -- the integer ranks of the reference -- at tensor-core speed.
 *   refine_capacity  worklist slots.  Every epilogue warp of the grid reserves slots in blocks of 128 (one global
 *                    atomic per block), so allow SMs x 8 x 128 (151,552 on a B200) on top of the expected band
 *                    (a few candidates per query); 2^19 + 128 t is a comfortable choice.
 *   refine_ws        blp_fast_refine_bytes(refine_capacity) bytes; the first 16 bytes are a header
 *                    {uint32 count; uint32 overflow; ...}: ZERO it once before the first call.  `overflow` is
 *                    sticky: non-zero after a call means more than refine_capacity candidates fell into the band
 *                    (e.g. a table of identical rows) and the counters of that call are NOT valid -- re-run the
 *                    chunk with blp_rank_sweep.  `count` holds the entries of the last call.
 *   ent, rel_weight, h_rows, t_rows must be 16-byte aligned.  Other arguments as blp_rank_sweep_fast. */
int64_t blp_fast_refine_bytes(int64_t refine_capacity);
int blp_rank_sweep_fast_exact(int model, const float *ent, int64_t n_local, int64_t ent_offset, int d,
                              const float *rel_weight, int64_t num_rel, const int64_t *triples, int64_t t,
                              const float *h_rows, const float *t_rows,
                              const int64_t *filt_indptr, const int64_t *filt_idx, int64_t tail_off,
                              int32_t *gt, int32_t *ge, int32_t *gt_f, int32_t *ge_f, float *true_score,
                              const void *table_ws, void *query_ws, void *refine_ws, int64_t refine_capacity,
                              void *stream);

/* utils.py:106-109 + train.py:154-157 in one launch: per-query reciprocal ranks / hits
 * (either may be NULL) and the fp64 accumulators of blp_metrics_reduce. */
int blp_rank_metrics(const int32_t *gt, const int32_t *ge, int64_t q, const int64_t *k_values_host, int nk,
                     float *recip, uint8_t *hits, double *sums, void *stream);

/* ---- a5-a8  LinkPrediction.compute_loss, forward + backward ---------------
 * Replaces models.py:51-70 and its autograd graph with one fused pass.
 *   ent_embs   [b, 2, d]     (models.py:56 chunk -> heads, tails)
 *   rel_weight [num_rel, d], rels [b] int64   (models.py:55 rel_emb lookup)
 *   neg_idx    int64, logical (b, k, 2) with element strides (s0, s1, s2): the
 *              reference sampler returns a transposed view (data.py:78-79);
 *              values index ent_embs.view(2b, d) (models.py:65)
 *   regularizer  models.py:59-62 (0 = off)
 *   loss_out   [1]   model_loss + reg_loss (models.py:70)
 *   pos_scores [b], neg_scores [b, k]   (written; neg_scores may be NULL)
 *   grad_ent [b,2,d], grad_rel_weight [num_rel,d]: d loss / d ent_embs and
 *              d loss / d rel_emb.weight (dense, like nn.Embedding's backward)
 *              for an upstream gradient of 1, or both NULL for forward only.
 *              Both are overwritten.  (The loss is linear in the upstream
 *              gradient, so autograd's backward is blp_scale by grad_out.)
 *   workspace  blp_train_workspace_bytes(b, k) bytes, zero-filled once by the
 *              caller; the call leaves it zero-filled again.  One workspace
 *              per concurrently running stream.
 * Conventions follow the reference ops: margin mask keeps gradient where
 * 1 - pos + neg == 0 (models.py:253), sign(0) = 0 for TransE, softplus
 * threshold 20 (models.py:258). */
int64_t blp_train_workspace_bytes(int64_t b, int64_t k);
int blp_train_loss(int model, int loss, const float *ent_embs, const float *rel_weight,
                   const int64_t *rels, int64_t num_rel,
                   const int64_t *neg_idx, int64_t s0, int64_t s1, int64_t s2,
                   int64_t b, int64_t k, int d, float regularizer,
                   float *loss_out, float *pos_scores, float *neg_scores,
                   float *grad_ent, float *grad_rel_weight, void *workspace, void *stream);

/* y[i] = x[i] * *scale_dev (scale_dev is a DEVICE scalar; y may alias x):
 * autograd's backward of the fused loss multiplies the stored unit-upstream
 * gradients by grad_output. */
int blp_scale(float *y, const float *x, const float *scale_dev, int64_t n, void *stream);

/* ---- a5, a6  loss_fn(pos_scores, neg_scores)  (models.py:251-258) ----------
 * Stand-alone form for direct callers of margin_loss / nll_loss on
 * materialised scores: pos [b], neg (b, k) with row stride neg_row_stride.
 * loss_out [1]; grad_pos [b] and grad_neg [b, k] (contiguous) receive
 * d loss / d pos and d loss / d neg, either may be NULL. */
int blp_pair_loss(int loss, const float *pos, const float *neg, int64_t neg_row_stride,
                  int64_t b, int64_t k, float *loss_out, float *grad_pos, float *grad_neg,
                  void *stream);
/* ---- a7  l2_regularization(heads, tails, rels)  (models.py:261-266) --------
 * out[0] = (mean(heads^2) + mean(tails^2) + mean(rels^2)) / 3 over contiguous
 * tensors of n_* elements. */
int blp_l2_regularization(const float *heads, int64_t n_heads, const float *tails, int64_t n_tails,
                          const float *rels, int64_t n_rels, float *out, void *stream);

/* ---- a12 / next row f1  filtered ranks from a device-resident filter index -----------------
 * Replaces utils.get_triple_filters (utils.py:46-83: per-batch dense (B, N) bool masks built with Python
 * loops over the graph's edges), the mask H2D and the second get_metrics pass (train.py:159-167).
 * The filtering graph is turned ONCE per evaluation into two sorted arrays of composite keys
 * ((rel * n_rows + fixed_row) * n_rows + other_row); the known tails of (head, rel) / heads of (tail, rel)
 * are one run found by binary search.
 *   edges      [num_edges, 3] int64 (head id, tail id, rel) -- graph.edges(keys=True) of the filtering graph
 *   ent2idx    [n_ids] int64 entity id -> table row or -1 (utils.py:31-43); NULL = ids are rows already
 *   n_rows     rows of the WHOLE entity table (all shards); num_rel relation count
 *   index_ws   blp_filter_index_bytes(num_edges) bytes, 256-byte aligned; holds the index afterwards
 * Edges touching an entity without a row are dropped (utils.py:72-73,79-80); parallel edges collapse. */
int64_t blp_filter_index_bytes(int64_t num_edges);
int blp_filter_index_build(const int64_t *edges, int64_t num_edges, const int64_t *ent2idx, int64_t n_ids,
                           int64_t n_rows, int64_t num_rel, void *index_ws, int64_t index_bytes, void *stream);
/* Filtered counters from the raw ones (run after blp_rank_sweep / blp_rank_sweep_fast on the same stream):
 * gt_f = gt - #{filtered candidates in this shard with s > s_true}, likewise ge_f.  A candidate is filtered
 * for a head query iff (cand, tail, rel) is an edge and cand != head (utils.py:76-81); for a tail query iff
 * (head, cand, rel) is an edge and cand != tail (utils.py:69-74).  triples / h_rows / t_rows / tail_off as
 * blp_rank_sweep; true_score, gt, ge are that call's outputs. */
int blp_filter_correct(int model, const float *ent, int64_t n_local, int64_t ent_offset, int d,
                       const float *rel_weight, int64_t num_rel, const int64_t *triples, int64_t t,
                       const float *h_rows, const float *t_rows, const void *index_ws, int64_t num_edges,
                       int64_t n_rows, int64_t tail_off, const float *true_score, const int32_t *gt,
                       const int32_t *ge, int32_t *gt_f, int32_t *ge_f, void *stream);

/* blp_filter_correct for the score-matrix path (train.py:141-171 left untouched): the lookup keys come from
 * `triples` (t, 3) = (head row, tail row, relation id); the operand rows are the dense (t, d) blocks the call site
 * gathered itself (head_embs, tail_embs, rel_embs of train.py:141-143). */
int blp_filter_correct_rows(int model, const float *ent, int64_t n_local, int64_t ent_offset, int d,
                            const int64_t *triples, int64_t t, const float *h_rows, const float *t_rows,
                            const float *r_rows, const void *index_ws, int64_t num_edges, int64_t n_rows,
                            int64_t num_rel, int64_t tail_off, const float *true_score, const int32_t *gt,
                            const int32_t *ge, int32_t *gt_f, int32_t *ge_f, void *stream);
/* train.py:164-167 from the reference's own dense bool mask (`pred[filter_mask] = pred.min() - 1.0` followed by
 * get_metrics): filtered counters of the n_hq + n_tq queries of blp_rank_queries (same query arguments and output
 * order).  mask: (n_hq + n_tq, ld_mask) bytes, column j = candidate row ent_offset + j; only the masked candidates
 * are re-scored.  A masked true candidate gets the reference's semantics (its score becomes min - 1). */
int blp_filter_correct_mask(int model, const float *ent, int64_t n_local, int64_t ent_offset, int d,
                            const float *hq_tails, const float *hq_rels, const int64_t *hq_true, int64_t n_hq,
                            const float *tq_heads, const float *tq_rels, const int64_t *tq_true, int64_t n_tq,
                            const uint8_t *mask, int64_t ld_mask, const float *true_score, const int32_t *gt,
                            const int32_t *ge, int32_t *gt_f, int32_t *ge_f, void *stream);

/* ---- next row f2  MRR breakdowns of train.py:173-188 (utils.py:114-168) in one launch ---------
 *   recip      reciprocal ranks, head query i at [i], tail query i at [tail_off + i]
 *   triples    [t, 3] int64 (head id, tail id, rel) ENTITY IDS (as the reference's loops see them)
 *   is_new     [n_ids] bytes, 1 = entity is new (utils.split_by_new_position), or NULL
 *   rel_categories [num_rel] int64 in [0, 4) (utils.split_by_category), or NULL
 *   out        18 doubles (overwritten): mrr_by_position[3] (both new / head new / tail new), counts[3],
 *              mrr_by_category[2][4] (row 0 head prediction, row 1 tail prediction), category counts[4] */
int blp_mrr_breakdown(const float *recip, int64_t t, int64_t tail_off, const int64_t *triples,
                      const uint8_t *is_new, int64_t n_ids, const int64_t *rel_categories, int64_t num_rel,
                      double *out, void *stream);

/* ---- a14 / next row f3  in-batch negative sampling indices (data.py:35-81) --------------------
 * out is the (num_neg, batch * repeats, 2) int64 buffer whose transposed view (batch * repeats, num_neg, 2),
 * element strides (2, 2 * batch * repeats, 1), is what get_negative_sampling_indices returns (data.py:77-79)
 * and what blp_train_loss reads in place.  Every negative keeps one entity of its own pair (2b or 2b + 1)
 * and replaces the other (fair coin, data.py:71) by a uniform draw from the 2 * batch - 2 entities of the
 * other pairs (data.py:60-65).  Philox4x32-10 keyed by seed, counter (element, offset): the stream differs
 * from torch's CPU generator by construction -- distributional parity only. */
int blp_negative_sample(int64_t batch, int64_t num_neg, int64_t repeats, uint64_t seed, uint64_t offset,
                        int64_t *out, void *stream);

/* ---- a13 / next row f4  entity-table production: F.normalize + scatter into a row shard ---------
 * Replaces `F.normalize(ent_emb, dim=-1)` (models.py:38-43, TransE only) and
 * `ent_emb[idx:idx + bs] = batch_emb` (train.py:95-123) for a row-sharded table: source row i of
 * emb [m, d] goes to global row dst_rows[i] (or row0 + i when dst_rows is NULL); rows outside
 * [ent_offset, ent_offset + n_local) are skipped, so every rank can be handed the same encoder batch
 * and keeps only what it owns.  normalize != 0: x / max(||x||_2, 1e-12) in ATen's CPU order (bit-equal).
 * d == 128, dst_rows == NULL and m >= 4096 (bulk table production): a persistent TMA pipeline (bulk loads, in-place
 * normalisation in shared memory, bulk stores) at ~0.91 of the HBM copy peak. */
int blp_store_rows(const float *emb, int64_t m, int d, int normalize, const int64_t *dst_rows, int64_t row0,
                   float *ent_shard, int64_t n_local, int64_t ent_offset, void *stream);

/* ---- measurement aid ------------------------------------------------------
 * FP32 pipe micro-benchmarks used by bench.py to measure the lane-op rate the
 * ALU-bound exact sweeps are compared against ("measure
 * with a microbenchmark").  Every thread of a full grid runs `iters` rounds of
 * 16 independent dependency chains of the chosen instruction mix and writes one
 * float to sink[n_threads]:
 *   0 = FADD            acc = acc + c
 *   1 = FADD |x|        acc = acc + |u - e|   (2 FADD, the TransE tail-pred step)
 *   2 = add.f32x2       two chains per instruction
 *   3 = f32x2 TransE    {acc0,acc1} += |{u0,u1} - {e,e}|  (2 packed adds + 2 LOP)
 *   4 = FMUL + FADD     acc = acc + v * e     (the DistMult tail-pred step, unfused)
 *   5 = f32x2 DistMult  {acc0,acc1} += {v0,v1} * {e,e}  (packed product as fma(a,b,-0) + packed add)
 * *lane_ops_host receives the number of fp32 lane operations issued (adds and
 * multiplies). */
int blp_pipe_probe(int variant, float *sink, int64_t n_threads, int iters, double *lane_ops_host, void *stream);
/* Measurement aid: fp32 reduction (red.global.add.v4.f32) throughput into an L2-resident [rows, 128] table with the
 * access pattern of the compute_loss gradient scatter (whole 512-byte rows at random): one CTA of 64 lane groups per
 * SM, `iters` rows per group; *bytes_host receives the bytes added.  The table is modified. */
int blp_atomic_probe(float *table, int64_t rows, int iters, double *bytes_host, void *stream);

/* Bracket the dominant kernel of the following calls on THIS thread with the
 * caller's CUDA events (cudaEvent_t passed as void*), recorded on the stream the
 * kernel is launched on: which = 1 the eval sweep kernel (blp_eval_rank /
 * blp_score_bcast fast path), 2 the fused train kernel, 0 switches it off.
 * bench.py uses this for the per-launch duration behind `roofline.achieved`. */
int blp_profile_events(int which, void *start_event, void *stop_event);

/* Number of kernels the last call on this thread launched (bench.py's
 * `gpu_launches` is counted from this). */
int blp_last_launch_count(void);

/* measurement aid: sweep launches of this host thread write 16 globaltimer (ns) slots per CTA into `buffer`
 * (>= 148 * 16 uint64; slots: 0 start, 1 setup done, 2 true scores, 3 operands folded, 4 first tile landed,
 * 5 tiles done, 6 ticket taken, 7 epilogue done, 8 #segments, 9 #work items); NULL switches it off. */
int blp_debug_timestamps(void *buffer);

#ifdef __cplusplus
}
#endif
#endif /* BLP_B200_H */
