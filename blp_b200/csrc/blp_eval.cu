// blp_eval.cu -- fused full-entity scoring + rank counting (a1-a4, a10-a12).
//
// Replaces train.py:141-171 + utils.py:86-111 of the reference: for every query
// (a test triple with its head or its tail removed) score ALL candidate entity
// rows and reduce the scores to two integers, gt = #{s_j > s_true} and
// ge = #{s_j >= s_true}.  Neither the (B, N, D) broadcast nor the (2B, N) score
// matrix is ever materialised.
//
// The d == 128 sweep kernel lives in blp_sweep.cu; this file holds the C-ABI entry points, the
// true-score / filter-correction / generic-width kernels and get_metrics.
#include "blp_sweep.h"

namespace blp {

// ---- true-triple scores + counter reset ------------------------------------
// s_true for head query i and tail query i is the same number: the score of
// (h_i, r_i, t_i) (pred.gather(true_idx), utils.py:103).  One warp per triple: the
// lanes stage the three rows in shared memory with coalesced loads, lane 0 then
// replays the reference's sequential / ATen-order reduction from there (the same
// score_exact code every other exact path uses, so all of them agree bit for bit).
constexpr int kTrueWarps = 4;
__global__ void __launch_bounds__(kTrueWarps * 32) true_score_kernel(int model, const RowRef hr, const RowRef tr,
                                                                    const RowRef rr, long long b, long long tail_off,
                                                                    int d, int staged, float *__restrict__ true_score,
                                                                    int *__restrict__ gt, int *__restrict__ ge) {
    extern __shared__ float ts_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long i = (long long)blockIdx.x * kTrueWarps + warp;
    if (i >= b) return;
    const float *h = hr.row(i, d), *t = tr.row(i, d), *r = rr.row(i, d);
    if (staged) {
        float *sh = ts_smem + (size_t)warp * 3 * d, *stt = sh + d, *sr = stt + d;
        for (int j = lane; j < d; j += 32) {
            sh[j] = h[j];
            stt[j] = t[j];
            sr[j] = r[j];
        }
        __syncwarp();
        h = sh; t = stt; r = sr;
    }
    if (lane == 0) {
        float s = score_exact_dyn(model, h, t, r, d);
        // an index outside the table (train.py:137-138 asserts this never happens): NaN compares false
        // against every candidate, so the query reports gt = ge = 0 and is detectable
        if (!(hr.in_range(i) && tr.in_range(i) && rr.in_range(i))) s = __int_as_float(0x7fc00000);
        true_score[i] = s;
        true_score[tail_off + i] = s;
        gt[i] = 0; gt[tail_off + i] = 0;
        ge[i] = 0; ge[tail_off + i] = 0;
    }
}

// d == 128 specialisation: one warp per triple.  The 128 (64) per-position terms are computed in parallel,
// one float4 per lane, and parked in shared memory; only the reference's summation order is replayed
// serially: 128 dependent adds for torch.norm(p=1) (lane 0), or ATen's 8-lane x 4-accumulator cascade for
// torch.sum (lanes 0-7 run their chains in parallel, lane 0 folds the 8 lane sums).  Same bits as
// score_exact, ~10x shorter critical path.
constexpr int kTrue128Warps = 8;
template <int MODEL>
__global__ void __launch_bounds__(kTrue128Warps * 32) true_score128_kernel(const RowRef hr, const RowRef tr, const RowRef rr,
                                                                           long long b, long long tail_off,
                                                                           float *__restrict__ true_score,
                                                                           int *__restrict__ gt, int *__restrict__ ge) {
    __shared__ __align__(16) float terms[kTrue128Warps][kD];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long i = (long long)blockIdx.x * kTrue128Warps + warp;
    if (i >= b) return;
    float s = true_score_warp128<MODEL>(hr.row(i, kD), tr.row(i, kD), rr.row(i, kD), terms[warp], lane);
    if (lane == 0) {
        if (!(hr.in_range(i) && tr.in_range(i) && rr.in_range(i))) s = __int_as_float(0x7fc00000);
        true_score[i] = s;
        true_score[tail_off + i] = s;
        gt[i] = 0; gt[tail_off + i] = 0;
        ge[i] = 0; ge[tail_off + i] = 0;
    }
}

// ---- filtered ranks: sparse correction (train.py:159-167) -------------------
// One warp per query: re-score the query's filtered candidates that live in this
// shard and remove their contribution from the raw counts.
__global__ void filter_correct_kernel(int model, const float *__restrict__ ent, long long n_local, long long ent_offset,
                                      int d, const RowRef hr, const RowRef tr, const RowRef rr, long long b,
                                      long long tail_off, const long long *__restrict__ indptr,
                                      const long long *__restrict__ idx, const float *__restrict__ true_score,
                                      const int *__restrict__ gt, const int *__restrict__ ge, int *__restrict__ gt_f,
                                      int *__restrict__ ge_f) {
    const long long q = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // CSR row: heads then tails
    const int lane = threadIdx.x & 31;
    if (q >= 2 * b) return;
    const bool head_pred = q < b;
    const long long i = head_pred ? q : q - b;
    const long long o = head_pred ? i : tail_off + i;                               // output slot
    const float st = true_score[o];
    int cg = 0, ce = 0;
    if (indptr) {
        const float *h = hr.row(i, d), *t = tr.row(i, d), *r = rr.row(i, d);
        for (long long p = indptr[q] + lane; p < indptr[q + 1]; p += 32) {
            const long long row = idx[p] - ent_offset;
            if (row < 0 || row >= n_local) continue;
            const float *e = ent + row * d;
            const float s = head_pred ? score_exact_dyn(model, e, t, r, d) : score_exact_dyn(model, h, e, r, d);
            cg += s > st;
            ce += s >= st;
        }
    }
    cg = __reduce_add_sync(0xffffffffu, cg);
    ce = __reduce_add_sync(0xffffffffu, ce);
    if (lane == 0) {
        gt_f[o] = gt[o] - cg;
        ge_f[o] = ge[o] - ce;
    }
}

// ---- generic-width sweep (d != 128): one thread per (query, candidate) -----
__global__ void sweep_generic_kernel(int model, const float *__restrict__ ent, long long n_local, int d, const RowRef hr,
                                     const RowRef tr, const RowRef rr, long long b, long long tail_off,
                                     const float *__restrict__ true_score, int *__restrict__ gt, int *__restrict__ ge) {
    for (long long q = blockIdx.y; q < 2 * b; q += gridDim.y) {
        const bool head_pred = q < b;
        const long long i = head_pred ? q : q - b;
        const long long o = head_pred ? i : tail_off + i;
        const float st = true_score[o];
        const float *h = hr.row(i, d), *t = tr.row(i, d), *r = rr.row(i, d);
        int cg = 0, ce = 0;
        for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n_local; c += (long long)gridDim.x * blockDim.x) {
            const float *e = ent + c * d;
            const float s = head_pred ? score_exact_dyn(model, e, t, r, d) : score_exact_dyn(model, h, e, r, d);
            cg += s > st;
            ce += s >= st;
        }
        cg = __reduce_add_sync(0xffffffffu, cg);
        ce = __reduce_add_sync(0xffffffffu, ce);
        if ((threadIdx.x & 31) == 0 && (cg | ce)) {
            atomicAdd(&gt[o], cg);
            atomicAdd(&ge[o], ce);
        }
    }
}

// ---- materialising score_fn (generic broadcast) -----------------------------
__global__ void score_bcast_kernel(int model, const float *__restrict__ h, long long hsA, long long hsC,
                                   const float *__restrict__ t, long long tsA, long long tsC,
                                   const float *__restrict__ r, long long rsA, long long rsC, long long A, long long C,
                                   int d, float *__restrict__ out) {
    const long long total = A * C;
    for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
        const long long a = o / C, c = o - a * C;
        out[o] = score_exact_dyn(model, h + a * hsA + c * hsC, t + a * tsA + c * tsC, r + a * rsA + c * rsC, d);
    }
}

// ---- get_metrics on a materialised score matrix (utils.py:103-105) ----------
__global__ void rank_counts_kernel(const float *__restrict__ pred, long long n, long long row_stride,
                                   const long long *__restrict__ true_idx, int *__restrict__ gt, int *__restrict__ ge) {
    const long long q = blockIdx.x;
    const float *row = pred + q * row_stride;
    const float st = row[true_idx[q]];
    int cg = 0, ce = 0;
    for (long long j = threadIdx.x; j < n; j += blockDim.x) {
        const float s = row[j];
        cg += s > st;
        ce += s >= st;
    }
    __shared__ int sg[32], se[32];
    cg = __reduce_add_sync(0xffffffffu, cg);
    ce = __reduce_add_sync(0xffffffffu, ce);
    if ((threadIdx.x & 31) == 0) { sg[threadIdx.x >> 5] = cg; se[threadIdx.x >> 5] = ce; }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int nw = blockDim.x >> 5;
        cg = threadIdx.x < nw ? sg[threadIdx.x] : 0;
        ce = threadIdx.x < nw ? se[threadIdx.x] : 0;
        cg = __reduce_add_sync(0xffffffffu, cg);
        ce = __reduce_add_sync(0xffffffffu, ce);
        if (threadIdx.x == 0) { gt[q] = cg; ge[q] = ce; }
    }
}

struct KValues { long long k[8]; int nk; };
// utils.py:106-109: avg = float(best + worst) * 0.5; 1 / avg; avg <= k
__global__ void metrics_kernel(const int *__restrict__ gt, const int *__restrict__ ge, long long q, KValues kv,
                               float *__restrict__ recip, unsigned char *__restrict__ hits) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= q) return;
    const float avg = fmul((float)((long long)gt[i] + 1 + (long long)ge[i]), 0.5f);
    recip[i] = __frcp_rn(avg);
    for (int j = 0; j < kv.nk; ++j) hits[i * kv.nk + j] = avg <= (float)kv.k[j];
}

// train.py:154-157 accumulators: sums[0] = sum of reciprocal ranks, sums[1 + j] = number of hits at k_j.
// One CTA, fp64 accumulation in a fixed order (deterministic).
__global__ void __launch_bounds__(1024) metrics_reduce_kernel(const int *__restrict__ gt, const int *__restrict__ ge,
                                                              long long q, KValues kv, float *__restrict__ recip,
                                                              unsigned char *__restrict__ hits, double *__restrict__ sums) {
    __shared__ double scratch[32][9];
    double acc[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) acc[j] = 0.0;
    for (long long i = threadIdx.x; i < q; i += blockDim.x) {
        const float avg = fmul((float)((long long)gt[i] + 1 + (long long)ge[i]), 0.5f);
        const float rr = __frcp_rn(avg);
        acc[0] += (double)rr;
        if (recip) recip[i] = rr;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < kv.nk) {
                const bool hit = avg <= (float)kv.k[j];
                if (hit) acc[1 + j] += 1.0;
                if (hits) hits[i * kv.nk + j] = hit;
            }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
        if (lane == 0) scratch[warp][j] = acc[j];
    }
    __syncthreads();
    if (threadIdx.x <= kv.nk) {
        double tot = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += scratch[w][threadIdx.x];
        sums[threadIdx.x] = tot;
    }
}

// ---- host side ----------------------------------------------------------------

static int check_model_dim(int model, int d) {
    if (model < 0 || model > 3) { set_error("unknown relational model id %d", model); return BLP_EINVAL; }
    if (d <= 0) { set_error("d must be positive (got %d)", d); return BLP_EDIM; }
    if ((model == BLP_MODEL_COMPLEX || model == BLP_MODEL_SIMPLE) && (d & 1)) {
        set_error("complex/simple need an even d (got %d)", d);
        return BLP_EDIM;
    }
    return BLP_OK;
}

static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace blp

using namespace blp;

// Shared body of blp_eval_rank / blp_rank_sweep: true scores (+ counter reset), the sweep, the filter correction.
static int rank_impl(int model, const float *ent, int64_t n_local, int64_t ent_offset, int d, const RowRef &h,
                     const RowRef &t, const RowRef &r, int64_t b, int64_t tail_off, const int64_t *filt_indptr,
                     const int64_t *filt_idx, int32_t *gt, int32_t *ge, int32_t *gt_f, int32_t *ge_f, float *true_score,
                     cudaStream_t st, const long long *triples = nullptr, const void *fast_table_ws = nullptr,
                     void *fast_query_ws = nullptr, float *fast_scores = nullptr, long long fast_ld = 0,
                     int phases = 3 /* bit 0: true scores + counter reset, bit 1: sweep (+ filter correction) */) {
    const bool rows_aligned = aligned16(h.base) && aligned16(t.base) && aligned16(r.base);
    // tensor-core mode folds the true-score computation into its query-folding kernel (one launch less)
    const bool fused_true = (phases & 1) && (phases & 2) && fast_table_ws && n_local > 0 && d == kD && rows_aligned;
    if (fused_true || !(phases & 1)) {
        // nothing to launch here
    } else if (d == kD && rows_aligned) {
        const unsigned blocks = (unsigned)((b + kTrue128Warps - 1) / kTrue128Warps);
        switch (model) {
        case BLP_MODEL_TRANSE: true_score128_kernel<BLP_MODEL_TRANSE><<<blocks, kTrue128Warps * 32, 0, st>>>(h, t, r, b, tail_off, true_score, gt, ge); break;
        case BLP_MODEL_DISTMULT: true_score128_kernel<BLP_MODEL_DISTMULT><<<blocks, kTrue128Warps * 32, 0, st>>>(h, t, r, b, tail_off, true_score, gt, ge); break;
        case BLP_MODEL_COMPLEX: true_score128_kernel<BLP_MODEL_COMPLEX><<<blocks, kTrue128Warps * 32, 0, st>>>(h, t, r, b, tail_off, true_score, gt, ge); break;
        default: true_score128_kernel<BLP_MODEL_SIMPLE><<<blocks, kTrue128Warps * 32, 0, st>>>(h, t, r, b, tail_off, true_score, gt, ge); break;
        }
        count_launch();
    } else {
        const size_t ts_smem = (size_t)kTrueWarps * 3 * d * sizeof(float);
        const int staged = ts_smem <= 48 * 1024;
        true_score_kernel<<<(unsigned)((b + kTrueWarps - 1) / kTrueWarps), kTrueWarps * 32, staged ? ts_smem : 0, st>>>(
            model, h, t, r, b, tail_off, d, staged, true_score, gt, ge);
        count_launch();
    }
    BLP_CUDA(cudaGetLastError());
    if (!(phases & 2)) return BLP_OK;

    if (n_local > 0 && fast_table_ws) {
        // tensor-core mode: scores as a split-FP16 contraction on tcgen05, same counters (blp_fast.cu)
        const int rc = launch_fast_sweep(model, n_local, ent_offset, h, t, r, triples, b, tail_off, true_score, gt, ge,
                                         fast_table_ws, fast_query_ws, fast_scores, fast_ld, fused_true, st);
        if (rc) return rc;
    } else if (n_local > 0) {
        if (d == kD && aligned16(ent)) {
            SweepArgs a{};
            a.ent = ent; a.n_local = n_local; a.h = h; a.t = t; a.r = r; a.b = b; a.tail_off = tail_off;
            a.true_score = true_score; a.gt = gt; a.ge = ge; a.scores_out = nullptr; a.ld_scores = 0;
            a.roles = 3; a.groups = 0; a.use_tma = sweep_env_use_tma(); a.negzero2 = kNegZero2;
            const int rc = launch_sweep_dyn(model, a, st);
            if (rc) return rc;
        } else {
            const int threads = 256;
            long long bx = (n_local + threads - 1) / threads;
            if (bx > 1024) bx = 1024;
            const long long by = 2 * b < 65535 ? 2 * b : 65535;
            dim3 grid((unsigned)bx, (unsigned)by);
            sweep_generic_kernel<<<grid, threads, 0, st>>>(model, ent, n_local, d, h, t, r, b, tail_off, true_score, gt, ge);
            count_launch();
            BLP_CUDA(cudaGetLastError());
        }
    }
    if (gt_f && ge_f) {
        const long long warps = 2 * b;
        const int threads = 128;
        const long long blocks = (warps * 32 + threads - 1) / threads;
        filter_correct_kernel<<<(unsigned)blocks, threads, 0, st>>>(model, ent, n_local, ent_offset, d, h, t, r, b, tail_off,
                                                                    (const long long *)filt_indptr,
                                                                    (const long long *)filt_idx, true_score, gt, ge, gt_f, ge_f);
        count_launch();
        BLP_CUDA(cudaGetLastError());
    }
    return BLP_OK;
}

static int check_rank_args(int model, int d, int64_t b, int64_t n_local, const void *ent, const int64_t *filt_indptr,
                           const int64_t *filt_idx, const void *gt, const void *ge, const void *gt_f, const void *ge_f,
                           const void *true_score) {
    int rc = check_model_dim(model, d);
    if (rc) return rc;
    if (b < 0 || n_local < 0) { set_error("negative size"); return BLP_EINVAL; }
    if (b > 0 && (!gt || !ge || !true_score || (n_local > 0 && !ent))) { set_error("null pointer argument"); return BLP_EINVAL; }
    if (filt_indptr && (!gt_f || !ge_f || !filt_idx)) { set_error("filters given without gt_f/ge_f/filt_idx"); return BLP_EINVAL; }
    if (n_local >= (1ll << 31)) { set_error("n_local too large for int32 counters"); return BLP_EINVAL; }
    return BLP_OK;
}

extern "C" int blp_eval_rank(int model, const float *ent, int64_t n_local, int64_t ent_offset, int d,
                             const float *h_rows, const float *t_rows, const float *r_rows, int64_t b,
                             const int64_t *filt_indptr, const int64_t *filt_idx, int32_t *gt, int32_t *ge,
                             int32_t *gt_f, int32_t *ge_f, float *true_score, void *stream) {
    reset_launch_count();
    int rc = check_rank_args(model, d, b, n_local, ent, filt_indptr, filt_idx, gt, ge, gt_f, ge_f, true_score);
    if (rc) return rc;
    if (b == 0) return BLP_OK;
    if (!h_rows || !t_rows || !r_rows) { set_error("null pointer argument"); return BLP_EINVAL; }
    return rank_impl(model, ent, n_local, ent_offset, d, dense_rows(h_rows), dense_rows(t_rows), dense_rows(r_rows), b, b,
                     filt_indptr, filt_idx, gt, ge, gt_f, ge_f, true_score, (cudaStream_t)stream);
}

extern "C" int blp_rank_sweep(int model, const float *ent, int64_t n_local, int64_t ent_offset, int d,
                              const float *rel_weight, int64_t num_rel, const int64_t *triples, int64_t t,
                              const float *h_rows, const float *t_rows, const int64_t *filt_indptr,
                              const int64_t *filt_idx, int64_t tail_off, int32_t *gt, int32_t *ge, int32_t *gt_f,
                              int32_t *ge_f, float *true_score, void *stream) {
    reset_launch_count();
    int rc = check_rank_args(model, d, t, n_local, ent, filt_indptr, filt_idx, gt, ge, gt_f, ge_f, true_score);
    if (rc) return rc;
    if (t == 0) return BLP_OK;
    if (!rel_weight || !triples || num_rel <= 0) { set_error("null pointer argument"); return BLP_EINVAL; }
    if ((h_rows == nullptr) != (t_rows == nullptr)) { set_error("h_rows and t_rows must both be given or both NULL"); return BLP_EINVAL; }
    if (tail_off < t) { set_error("tail_off must be >= t"); return BLP_EINVAL; }
    const long long *tr = (const long long *)triples;
    // train.py:141-143: head_embs = ent_emb[heads], tail_embs = ent_emb[tails], rel_embs = rel_emb(rels)
    const RowRef h = h_rows ? dense_rows(h_rows) : RowRef{ent, tr + 0, 3, ent_offset, n_local};
    const RowRef tt = t_rows ? dense_rows(t_rows) : RowRef{ent, tr + 1, 3, ent_offset, n_local};
    const RowRef r = RowRef{rel_weight, tr + 2, 3, 0, num_rel};
    if (!h_rows && n_local == 0) { set_error("cannot gather query rows from an empty shard; pass h_rows / t_rows"); return BLP_EINVAL; }
    return rank_impl(model, ent, n_local, ent_offset, d, h, tt, r, t, tail_off, filt_indptr, filt_idx, gt, ge, gt_f, ge_f,
                     true_score, (cudaStream_t)stream);
}

// Chunked sweeps (the reference's eval batches): the true scores and the counter reset of ALL t triples in one
// launch, then one sweep launch per chunk (blp_rank_sweep_counts) -- per chunk this halves the launches, which is
// what bounds the entity-sharded Wikidata5M-scale sweep at eval batch 2 (49 us of HBM time per batch on 8 GPUs).
extern "C" int blp_true_scores(int model, const float *ent, int64_t n_local, int64_t ent_offset, int d,
                               const float *rel_weight, int64_t num_rel, const int64_t *triples, int64_t t,
                               const float *h_rows, const float *t_rows, int64_t tail_off, int32_t *gt, int32_t *ge,
                               float *true_score, void *stream) {
    reset_launch_count();
    int rc = check_rank_args(model, d, t, n_local, ent, nullptr, nullptr, gt, ge, nullptr, nullptr, true_score);
    if (rc) return rc;
    if (t == 0) return BLP_OK;
    if (!rel_weight || !triples || num_rel <= 0) { set_error("null pointer argument"); return BLP_EINVAL; }
    if ((h_rows == nullptr) != (t_rows == nullptr)) { set_error("h_rows and t_rows must both be given or both NULL"); return BLP_EINVAL; }
    if (tail_off < t) { set_error("tail_off must be >= t"); return BLP_EINVAL; }
    if (!h_rows && n_local == 0) { set_error("cannot gather query rows from an empty shard; pass h_rows / t_rows"); return BLP_EINVAL; }
    const long long *tr = (const long long *)triples;
    const RowRef h = h_rows ? dense_rows(h_rows) : RowRef{ent, tr + 0, 3, ent_offset, n_local};
    const RowRef tt = t_rows ? dense_rows(t_rows) : RowRef{ent, tr + 1, 3, ent_offset, n_local};
    const RowRef r = RowRef{rel_weight, tr + 2, 3, 0, num_rel};
    return rank_impl(model, ent, n_local, ent_offset, d, h, tt, r, t, tail_off, nullptr, nullptr, gt, ge, nullptr, nullptr,
                     true_score, (cudaStream_t)stream, nullptr, nullptr, nullptr, nullptr, 0, 1);
}

extern "C" int blp_rank_sweep_counts(int model, const float *ent, int64_t n_local, int64_t ent_offset, int d,
                                     const float *rel_weight, int64_t num_rel, const int64_t *triples, int64_t t,
                                     const float *h_rows, const float *t_rows, const int64_t *filt_indptr,
                                     const int64_t *filt_idx, int64_t tail_off, int32_t *gt, int32_t *ge, int32_t *gt_f,
                                     int32_t *ge_f, const float *true_score, void *stream) {
    reset_launch_count();
    int rc = check_rank_args(model, d, t, n_local, ent, filt_indptr, filt_idx, gt, ge, gt_f, ge_f, true_score);
    if (rc) return rc;
    if (t == 0) return BLP_OK;
    if (!rel_weight || !triples || num_rel <= 0) { set_error("null pointer argument"); return BLP_EINVAL; }
    if ((h_rows == nullptr) != (t_rows == nullptr)) { set_error("h_rows and t_rows must both be given or both NULL"); return BLP_EINVAL; }
    if (tail_off < t) { set_error("tail_off must be >= t"); return BLP_EINVAL; }
    if (!h_rows && n_local == 0) { set_error("cannot gather query rows from an empty shard; pass h_rows / t_rows"); return BLP_EINVAL; }
    const long long *tr = (const long long *)triples;
    const RowRef h = h_rows ? dense_rows(h_rows) : RowRef{ent, tr + 0, 3, ent_offset, n_local};
    const RowRef tt = t_rows ? dense_rows(t_rows) : RowRef{ent, tr + 1, 3, ent_offset, n_local};
    const RowRef r = RowRef{rel_weight, tr + 2, 3, 0, num_rel};
    return rank_impl(model, ent, n_local, ent_offset, d, h, tt, r, t, tail_off, filt_indptr, filt_idx, gt, ge, gt_f, ge_f,
                     const_cast<float *>(true_score), (cudaStream_t)stream, nullptr, nullptr, nullptr, nullptr, 0, 2);
}

extern "C" int64_t blp_fast_table_bytes(int64_t n_local) { return fast_table_ws_bytes(n_local); }
extern "C" int64_t blp_fast_query_bytes(int64_t t) { return fast_query_ws_bytes(t); }

extern "C" int blp_fast_prepare_table(const float *ent, int64_t n_local, int d, void *table_ws, void *stream) {
    reset_launch_count();
    if (d != kD) { set_error("fast mode is specialised for d = %d (got %d)", kD, d); return BLP_EDIM; }
    if (n_local < 0 || !table_ws || (n_local > 0 && !ent)) { set_error("bad argument"); return BLP_EINVAL; }
    if (!aligned16(ent) || !aligned16(table_ws)) { set_error("ent / table_ws must be 16-byte aligned"); return BLP_EINVAL; }
    return fast_prepare_table(ent, n_local, table_ws, (cudaStream_t)stream);
}

extern "C" int blp_rank_sweep_fast(int model, const float *ent, int64_t n_local, int64_t ent_offset, int d,
                                   const float *rel_weight, int64_t num_rel, const int64_t *triples, int64_t t,
                                   const float *h_rows, const float *t_rows, const int64_t *filt_indptr,
                                   const int64_t *filt_idx, int64_t tail_off, int32_t *gt, int32_t *ge, int32_t *gt_f,
                                   int32_t *ge_f, float *true_score, const void *table_ws, void *query_ws,
                                   float *scores_out, int64_t ld_scores, void *stream) {
    reset_launch_count();
    int rc = check_rank_args(model, d, t, n_local, ent, filt_indptr, filt_idx, gt, ge, gt_f, ge_f, true_score);
    if (rc) return rc;
    if (d != kD) { set_error("fast mode is specialised for d = %d (got %d)", kD, d); return BLP_EDIM; }
    if (model == BLP_MODEL_TRANSE) { set_error("fast (tensor-core) mode covers distmult / complex / simple only"); return BLP_EINVAL; }
    if (t == 0) return BLP_OK;
    if (!rel_weight || !triples || num_rel <= 0 || !table_ws || !query_ws) { set_error("null pointer argument"); return BLP_EINVAL; }
    if (!aligned16(table_ws) || !aligned16(query_ws)) { set_error("workspaces must be 16-byte aligned"); return BLP_EINVAL; }
    if ((h_rows == nullptr) != (t_rows == nullptr)) { set_error("h_rows and t_rows must both be given or both NULL"); return BLP_EINVAL; }
    if (tail_off < t) { set_error("tail_off must be >= t"); return BLP_EINVAL; }
    if (scores_out && ld_scores < n_local) { set_error("ld_scores must be >= n_local"); return BLP_EINVAL; }
    const long long *tr = (const long long *)triples;
    const RowRef h = h_rows ? dense_rows(h_rows) : RowRef{ent, tr + 0, 3, ent_offset, n_local};
    const RowRef tt = t_rows ? dense_rows(t_rows) : RowRef{ent, tr + 1, 3, ent_offset, n_local};
    const RowRef r = RowRef{rel_weight, tr + 2, 3, 0, num_rel};
    if (!h_rows && n_local == 0) { set_error("cannot gather query rows from an empty shard; pass h_rows / t_rows"); return BLP_EINVAL; }
    return rank_impl(model, ent, n_local, ent_offset, d, h, tt, r, t, tail_off, filt_indptr, filt_idx, gt, ge, gt_f, ge_f,
                     true_score, (cudaStream_t)stream, tr, table_ws, query_ws, scores_out, ld_scores);
}

extern "C" int blp_score_bcast(int model, const float *heads, int64_t hsA, int64_t hsC, const float *tails,
                               int64_t tsA, int64_t tsC, const float *rels, int64_t rsA, int64_t rsC, int64_t A,
                               int64_t C, int d, float *out, void *stream) {
    reset_launch_count();
    int rc = check_model_dim(model, d);
    if (rc) return rc;
    if (A < 0 || C < 0) { set_error("negative size"); return BLP_EINVAL; }
    if (A == 0 || C == 0) return BLP_OK;
    if (!heads || !tails || !rels || !out) { set_error("null pointer argument"); return BLP_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;

    // eval-shaped broadcast (train.py:146-147): one operand is the (1, C, d) table, the other two are (A, 1, d)
    const bool cand_h = hsA == 0 && hsC == d && tsC == 0 && rsC == 0 && tsA == d && rsA == d;
    const bool cand_t = tsA == 0 && tsC == d && hsC == 0 && rsC == 0 && hsA == d && rsA == d;
    if (d == kD && (cand_h || cand_t) && C >= 32 && aligned16(heads) && aligned16(tails) && aligned16(rels)) {
        SweepArgs a{};
        a.ent = cand_h ? heads : tails; a.n_local = C;
        // the unused query operand aliases a valid row block so the TMA staging reads defined memory
        a.h = dense_rows(cand_h ? tails : heads); a.t = dense_rows(cand_h ? tails : heads); a.r = dense_rows(rels);
        a.b = A; a.tail_off = A; a.true_score = nullptr; a.gt = nullptr; a.ge = nullptr; a.scores_out = out; a.ld_scores = C;
        a.roles = cand_h ? 1 : 2; a.groups = 0; a.use_tma = sweep_env_use_tma(); a.negzero2 = kNegZero2;
        return launch_sweep_dyn(model, a, st);
    }
    const long long total = A * C;
    long long blocks = (total + 127) / 128;
    if (blocks > 148 * 32) blocks = 148 * 32;
    score_bcast_kernel<<<(unsigned)blocks, 128, 0, st>>>(model, heads, hsA, hsC, tails, tsA, tsC, rels, rsA, rsC, A, C, d, out);
    count_launch();
    BLP_CUDA(cudaGetLastError());
    return BLP_OK;
}

extern "C" int blp_rank_counts(const float *pred, int64_t q, int64_t n, int64_t row_stride, const int64_t *true_idx,
                               int32_t *gt, int32_t *ge, void *stream) {
    reset_launch_count();
    if (q < 0 || n <= 0 || row_stride < n) { set_error("bad shape q=%lld n=%lld stride=%lld", (long long)q, (long long)n, (long long)row_stride); return BLP_EINVAL; }
    if (q == 0) return BLP_OK;
    if (!pred || !true_idx || !gt || !ge) { set_error("null pointer argument"); return BLP_EINVAL; }
    if (n >= (1ll << 31)) { set_error("n too large for int32 counters"); return BLP_EINVAL; }
    rank_counts_kernel<<<(unsigned)q, 256, 0, (cudaStream_t)stream>>>(pred, n, row_stride, (const long long *)true_idx, gt, ge);
    count_launch();
    BLP_CUDA(cudaGetLastError());
    return BLP_OK;
}

extern "C" int blp_metrics_from_counts(const int32_t *gt, const int32_t *ge, int64_t q, const int64_t *k_values_host,
                                       int nk, float *recip, uint8_t *hits, void *stream) {
    reset_launch_count();
    if (q < 0 || nk < 0 || nk > 8) { set_error("bad q or nk (nk <= 8)"); return BLP_EINVAL; }
    if (q == 0) return BLP_OK;
    if (!gt || !ge || !recip || (nk > 0 && (!hits || !k_values_host))) { set_error("null pointer argument"); return BLP_EINVAL; }
    KValues kv{};
    kv.nk = nk;
    for (int i = 0; i < nk; ++i) kv.k[i] = k_values_host[i];
    metrics_kernel<<<(unsigned)((q + 255) / 256), 256, 0, (cudaStream_t)stream>>>(gt, ge, q, kv, recip, hits);
    count_launch();
    BLP_CUDA(cudaGetLastError());
    return BLP_OK;
}

extern "C" int blp_metrics_reduce(const int32_t *gt, const int32_t *ge, int64_t q, const int64_t *k_values_host, int nk,
                                  double *sums, void *stream) {
    return blp_rank_metrics(gt, ge, q, k_values_host, nk, nullptr, nullptr, sums, stream);
}

extern "C" int blp_rank_metrics(const int32_t *gt, const int32_t *ge, int64_t q, const int64_t *k_values_host, int nk,
                                float *recip, uint8_t *hits, double *sums, void *stream) {
    reset_launch_count();
    if (q < 0 || nk < 0 || nk > 8) { set_error("bad q or nk (nk <= 8)"); return BLP_EINVAL; }
    if (!sums || (q > 0 && (!gt || !ge)) || (nk > 0 && !k_values_host)) { set_error("null pointer argument"); return BLP_EINVAL; }
    KValues kv{};
    kv.nk = nk;
    for (int i = 0; i < nk; ++i) kv.k[i] = k_values_host[i];
    metrics_reduce_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(gt, ge, q, kv, recip, hits, sums);
    count_launch();
    BLP_CUDA(cudaGetLastError());
    return BLP_OK;
}
