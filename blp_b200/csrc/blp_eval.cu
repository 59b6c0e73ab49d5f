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
#include <cooperative_groups.h>

#include "blp_sweep.h"

namespace cg = cooperative_groups;

namespace blp {

// ---- true-triple scores + counter reset ------------------------------------
// s_true for head query i and tail query i is the same number: the score of
// (h_i, r_i, t_i) (pred.gather(true_idx), utils.py:103).  One warp per triple: the
// lanes stage the three rows in shared memory with coalesced loads, lane 0 then
// replays the reference's sequential / ATen-order reduction from there (the same
// score_exact code every other exact path uses, so all of them agree bit for bit).
constexpr int kTrueWarps = 4;
// `split`: head query i and tail query i are unrelated queries with their own operand rows (hr2, tr2, rr2 for the
// tail queries); otherwise both queries of triple i share one true score.
__global__ void __launch_bounds__(kTrueWarps * 32) true_score_kernel(int model, const RowRef hr, const RowRef tr,
                                                                    const RowRef rr, const RowRef hr2, const RowRef tr2,
                                                                    const RowRef rr2, int split, long long b,
                                                                    long long tail_off, int d, int staged,
                                                                    float *__restrict__ true_score,
                                                                    int *__restrict__ gt, int *__restrict__ ge) {
    extern __shared__ float ts_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long j = (long long)blockIdx.x * kTrueWarps + warp;
    if (j >= (split ? 2 * b : b)) return;
    const bool second = j >= b;
    const long long i = second ? j - b : j;
    const RowRef &H = second ? hr2 : hr, &T = second ? tr2 : tr, &R = second ? rr2 : rr;
    const float *h = H.row(i, d), *t = T.row(i, d), *r = R.row(i, d);
    if (staged) {
        float *sh = ts_smem + (size_t)warp * 3 * d, *stt = sh + d, *sr = stt + d;
        for (int jj = lane; jj < d; jj += 32) {
            sh[jj] = h[jj];
            stt[jj] = t[jj];
            sr[jj] = r[jj];
        }
        __syncwarp();
        h = sh; t = stt; r = sr;
    }
    if (lane == 0) {
        float s = score_exact_dyn(model, h, t, r, d);
        // an index outside the table (train.py:137-138 asserts this never happens): NaN compares false
        // against every candidate, so the query reports gt = ge = 0 and is detectable
        if (!(H.in_range(i) && T.in_range(i) && R.in_range(i))) s = __int_as_float(0x7fc00000);
        if (!split || !second) { true_score[i] = s; gt[i] = 0; ge[i] = 0; }
        if (!split || second) { true_score[tail_off + i] = s; gt[tail_off + i] = 0; ge[tail_off + i] = 0; }
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
                                                                           const RowRef hr2, const RowRef tr2, const RowRef rr2,
                                                                           int split, long long b, long long tail_off,
                                                                           float *__restrict__ true_score,
                                                                           int *__restrict__ gt, int *__restrict__ ge) {
    __shared__ __align__(16) float terms[kTrue128Warps][kD];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long j = (long long)blockIdx.x * kTrue128Warps + warp;
    if (j >= (split ? 2 * b : b)) return;
    const bool second = j >= b;
    const long long i = second ? j - b : j;
    const RowRef &H = second ? hr2 : hr, &T = second ? tr2 : tr, &R = second ? rr2 : rr;
    float s = true_score_warp128<MODEL>(H.row(i, kD), T.row(i, kD), R.row(i, kD), terms[warp], lane);
    if (lane == 0) {
        if (!(H.in_range(i) && T.in_range(i) && R.in_range(i))) s = __int_as_float(0x7fc00000);
        if (!split || !second) { true_score[i] = s; gt[i] = 0; ge[i] = 0; }
        if (!split || second) { true_score[tail_off + i] = s; gt[tail_off + i] = 0; ge[tail_off + i] = 0; }
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
                                     const RowRef tr, const RowRef rr, const RowRef hr2, const RowRef tr2, const RowRef rr2,
                                     long long b, long long tail_off,
                                     const float *__restrict__ true_score, int *__restrict__ gt, int *__restrict__ ge) {
    for (long long q = blockIdx.y; q < 2 * b; q += gridDim.y) {
        const bool head_pred = q < b;
        const long long i = head_pred ? q : q - b;
        const long long o = head_pred ? i : tail_off + i;
        const float st = true_score[o];
        const float *h = (head_pred ? hr : hr2).row(i, d), *t = (head_pred ? tr : tr2).row(i, d), *r = (head_pred ? rr : rr2).row(i, d);
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
    // torch.gather raises on an out-of-range index (utils.py:103); a device kernel cannot, so the query is flagged
    // instead: NaN compares false against every score -> gt = ge = 0 (reciprocal rank 2.0, an impossible value)
    const long long ti = true_idx[q];
    const float st = (ti >= 0 && ti < n) ? row[ti] : __int_as_float(0x7fc00000);
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

// Whole evaluation sets (>= 4,096 queries, NK = kv.nk hit positions known at compile time, 16-byte aligned arrays): ONE
// thread-block cluster of 8 CTAs instead of one CTA -- the single-CTA kernel was bound by the instruction rate of one SM
// (40,960 queries: 34 us).  A thread takes four queries per iteration with 16-byte accesses, the 4 * NK hit bytes go out
// as NK packed words, hit counts stay integers; the per-CTA totals are folded by CTA 0 through distributed shared
// memory in rank order, so the fp64 sums are still deterministic (no atomics, no workspace).
constexpr int kMetricsCluster = 8;
template <int NK>
__global__ void __cluster_dims__(kMetricsCluster, 1, 1) __launch_bounds__(1024)
    metrics_reduce_cluster_kernel(const int *__restrict__ gt, const int *__restrict__ ge, long long q, KValues kv,
                                  float *__restrict__ recip, unsigned char *__restrict__ hits, double *__restrict__ sums) {
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned int rank = cluster.block_rank();
    __shared__ double scratch[32][1 + NK];
    __shared__ double block_tot[1 + NK];
    double rsum = 0.0;
    int cnt[NK];
    float kf[NK];
#pragma unroll
    for (int j = 0; j < NK; ++j) { cnt[j] = 0; kf[j] = (float)kv.k[j]; }
    const long long q4 = q / 4;
#pragma unroll 2
    for (long long i4 = (long long)rank * blockDim.x + threadIdx.x; i4 < q4; i4 += (long long)kMetricsCluster * blockDim.x) {
        const int4 g4 = reinterpret_cast<const int4 *>(gt)[i4], e4 = reinterpret_cast<const int4 *>(ge)[i4];
        const int g[4] = {g4.x, g4.y, g4.z, g4.w}, e[4] = {e4.x, e4.y, e4.z, e4.w};
        float rr[4];
        unsigned int hw[NK];
#pragma unroll
        for (int j = 0; j < NK; ++j) hw[j] = 0u;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float avg = fmul((float)((long long)g[u] + 1 + (long long)e[u]), 0.5f);
            rr[u] = __frcp_rn(avg);
            rsum += (double)rr[u];
#pragma unroll
            for (int j = 0; j < NK; ++j) {
                const unsigned int hit = avg <= kf[j] ? 1u : 0u;
                cnt[j] += (int)hit;
                const int p = u * NK + j;                       // byte p of the 4 * NK hit bytes of these four queries
                hw[p >> 2] |= hit << (8 * (p & 3));
            }
        }
        if (recip) reinterpret_cast<float4 *>(recip)[i4] = make_float4(rr[0], rr[1], rr[2], rr[3]);
        if (hits) {
            unsigned int *hp = reinterpret_cast<unsigned int *>(hits) + i4 * NK;
#pragma unroll
            for (int j = 0; j < NK; ++j) hp[j] = hw[j];
        }
    }
    if (rank == 0 && 4 * q4 + threadIdx.x < q) {          // the last q mod 4 queries
        const long long i = 4 * q4 + threadIdx.x;
        const float avg = fmul((float)((long long)gt[i] + 1 + (long long)ge[i]), 0.5f);
        const float rr = __frcp_rn(avg);
        rsum += (double)rr;
        if (recip) recip[i] = rr;
#pragma unroll
        for (int j = 0; j < NK; ++j) {
            const bool hit = avg <= kf[j];
            cnt[j] += hit ? 1 : 0;
            if (hits) hits[i * NK + j] = hit;
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rsum += __shfl_xor_sync(0xffffffffu, rsum, o);
    if (lane == 0) scratch[warp][0] = rsum;
#pragma unroll
    for (int j = 0; j < NK; ++j) {
        const int c = __reduce_add_sync(0xffffffffu, cnt[j]);
        if (lane == 0) scratch[warp][1 + j] = (double)c;
    }
    __syncthreads();
    if (threadIdx.x <= NK) {
        double tot = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += scratch[w][threadIdx.x];
        block_tot[threadIdx.x] = tot;
    }
    cluster.sync();
    if (rank == 0 && threadIdx.x <= NK) {
        double tot = 0.0;
        for (unsigned int r = 0; r < (unsigned int)kMetricsCluster; ++r) tot += *cluster.map_shared_rank(&block_tot[threadIdx.x], r);
        sums[threadIdx.x] = tot;
    }
    cluster.sync();                                       // the peers' shared memory stays valid until CTA 0 has read it
}

static void launch_metrics_reduce(const int *gt, const int *ge, long long q, const KValues &kv, float *recip, unsigned char *hits,
                                  double *sums, cudaStream_t st) {
    const bool al = ((reinterpret_cast<uintptr_t>(gt) | reinterpret_cast<uintptr_t>(ge) | reinterpret_cast<uintptr_t>(recip) |
                      reinterpret_cast<uintptr_t>(hits)) & 15u) == 0;
    // utils.py:86-111 is called with hit positions [1, 3, 10] everywhere in the reference (train.py:121); 4 covers (1, 3, 10, 100)
    if (al && q >= 4096 && kv.nk == 3) metrics_reduce_cluster_kernel<3><<<kMetricsCluster, 1024, 0, st>>>(gt, ge, q, kv, recip, hits, sums);
    else if (al && q >= 4096 && kv.nk == 4) metrics_reduce_cluster_kernel<4><<<kMetricsCluster, 1024, 0, st>>>(gt, ge, q, kv, recip, hits, sums);
    else metrics_reduce_kernel<<<1, 1024, 0, st>>>(gt, ge, q, kv, recip, hits, sums);
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

// One ranking problem: `b` head-prediction queries (operand rows h, t, r; the candidate plays `heads`, train.py:146)
// and `b` tail-prediction queries (rows h2, t2, r2; train.py:147).  In the reference's eval loop both come from the
// same b triples (split == 0: h2 == h, ...); the score-matrix path ranks unrelated queries (split == 1).
struct RankJob {
    int model;
    const float *ent; long long n_local, ent_offset; int d;
    RowRef h, t, r, h2, t2, r2; int split;
    long long b, tail_off;
    const long long *filt_indptr, *filt_idx;      // legacy CSR filter lists (triple mode only)
    int *gt, *ge, *gt_f, *ge_f; float *true_score;
    int roles;                                     // 3 both, 1 head queries only, 2 tail queries only (d == 128 only)
    int force_cfg;                                 // -1 auto
    // tensor-core mode
    const long long *triples; const void *fast_table_ws; void *fast_query_ws; float *fast_scores; long long fast_ld;
    void *fast_refine_ws; long long fast_refine_cap;   // filter + refine: exact ranks on the tensor path (NULL = plain fast mode)
    int phases;                                    // bit 0: true scores + counter reset, bit 1: sweep (+ CSR filter correction)
    int overlap;                                   // fused step only: programmatic dependent launch (blp_plan_set_overlap)
};

static RankJob make_job(int model, const float *ent, long long n_local, long long ent_offset, int d, const RowRef &h,
                        const RowRef &t, const RowRef &r, long long b, long long tail_off, int *gt, int *ge,
                        float *true_score) {
    RankJob j{};
    j.model = model; j.ent = ent; j.n_local = n_local; j.ent_offset = ent_offset; j.d = d;
    j.h = h; j.t = t; j.r = r; j.h2 = h; j.t2 = t; j.r2 = r; j.split = 0;
    j.b = b; j.tail_off = tail_off; j.gt = gt; j.ge = ge; j.true_score = true_score;
    j.roles = 3; j.force_cfg = -1; j.phases = 3;
    return j;
}

static void fill_sweep_args(SweepArgs &a, const RankJob &j) {
    a.ent = j.ent; a.n_local = j.n_local; a.h = j.h; a.t = j.t; a.r = j.r; a.h2 = j.h2; a.t2 = j.t2; a.r2 = j.r2;
    a.split_sets = j.split; a.b = j.b; a.tail_off = j.tail_off; a.true_score = j.true_score; a.gt = j.gt; a.ge = j.ge;
    a.scores_out = nullptr; a.ld_scores = 0; a.roles = j.roles; a.groups = 0; a.use_tma = sweep_env_use_tma();
    a.force_cfg = j.force_cfg; a.negzero2 = kNegZero2;
}

static bool job_rows_aligned(const RankJob &j) {
    return aligned16(j.h.base) && aligned16(j.t.base) && aligned16(j.r.base) && aligned16(j.h2.base) &&
           aligned16(j.t2.base) && aligned16(j.r2.base);
}

// Shared body of blp_eval_rank / blp_rank_sweep: true scores (+ counter reset), the sweep, the filter correction.
static int rank_impl(const RankJob &j, cudaStream_t st) {
    const int model = j.model, d = j.d;
    const long long b = j.b, tail_off = j.tail_off, n_local = j.n_local;
    const bool rows_aligned = job_rows_aligned(j);
    // tensor-core mode folds the true-score computation into its query-folding kernel (one launch less)
    const bool fused_true = (j.phases & 1) && (j.phases & 2) && j.fast_table_ws && n_local > 0 && d == kD && rows_aligned;
    const long long jobs = j.split ? 2 * b : b;
    if (fused_true || !(j.phases & 1)) {
        // nothing to launch here
    } else if (d == kD && rows_aligned) {
        const unsigned blocks = (unsigned)((jobs + kTrue128Warps - 1) / kTrue128Warps);
#define BLP_TS128(M) true_score128_kernel<M><<<blocks, kTrue128Warps * 32, 0, st>>>(j.h, j.t, j.r, j.h2, j.t2, j.r2, j.split, b, tail_off, j.true_score, j.gt, j.ge)
        switch (model) {
        case BLP_MODEL_TRANSE: BLP_TS128(BLP_MODEL_TRANSE); break;
        case BLP_MODEL_DISTMULT: BLP_TS128(BLP_MODEL_DISTMULT); break;
        case BLP_MODEL_COMPLEX: BLP_TS128(BLP_MODEL_COMPLEX); break;
        default: BLP_TS128(BLP_MODEL_SIMPLE); break;
        }
#undef BLP_TS128
        count_launch();
    } else {
        const size_t ts_smem = (size_t)kTrueWarps * 3 * d * sizeof(float);
        const int staged = ts_smem <= 48 * 1024;
        true_score_kernel<<<(unsigned)((jobs + kTrueWarps - 1) / kTrueWarps), kTrueWarps * 32, staged ? ts_smem : 0, st>>>(
            model, j.h, j.t, j.r, j.h2, j.t2, j.r2, j.split, b, tail_off, d, staged, j.true_score, j.gt, j.ge);
        count_launch();
    }
    BLP_CUDA(cudaGetLastError());
    if (!(j.phases & 2)) return BLP_OK;

    if (n_local > 0 && j.fast_table_ws) {
        // tensor-core mode: scores as a split-FP16 contraction on tcgen05, same counters (blp_fast.cu)
        const int rc = launch_fast_sweep(model, n_local, j.ent_offset, j.h, j.t, j.r, j.triples, b, tail_off, j.true_score,
                                         j.gt, j.ge, j.fast_table_ws, j.fast_query_ws, j.fast_scores, j.fast_ld, fused_true, j.ent,
                                         j.fast_refine_ws, j.fast_refine_cap, st);
        if (rc) return rc;
    } else if (n_local > 0) {
        if (d == kD && aligned16(j.ent)) {
            SweepArgs a{};
            fill_sweep_args(a, j);
            const int rc = launch_sweep_dyn(model, a, st);
            if (rc) return rc;
        } else if (sweep_wide_supports(model, d, j.ent) && job_rows_aligned(j) && j.roles == 3) {
            // TransE at the BOW widths (d = 300 / 768, ...): TMA-tiled, the row streams through in 64-float stages
            SweepArgs a{};
            fill_sweep_args(a, j);
            const int rc = launch_sweep_wide(a, d, st);
            if (rc) return rc;
        } else {
            const int threads = 256;
            long long bx = (n_local + threads - 1) / threads;
            if (bx > 1024) bx = 1024;
            const long long by = 2 * b < 65535 ? 2 * b : 65535;
            dim3 grid((unsigned)bx, (unsigned)by);
            sweep_generic_kernel<<<grid, threads, 0, st>>>(model, j.ent, n_local, d, j.h, j.t, j.r, j.h2, j.t2, j.r2, b, tail_off,
                                                           j.true_score, j.gt, j.ge);
            count_launch();
            BLP_CUDA(cudaGetLastError());
        }
    }
    if (j.gt_f && j.ge_f) {
        const long long warps = 2 * b;
        const int threads = 128;
        const long long blocks = (warps * 32 + threads - 1) / threads;
        filter_correct_kernel<<<(unsigned)blocks, threads, 0, st>>>(model, j.ent, n_local, j.ent_offset, d, j.h, j.t, j.r, b, tail_off,
                                                                    j.filt_indptr, j.filt_idx, j.true_score, j.gt, j.ge, j.gt_f, j.ge_f);
        count_launch();
        BLP_CUDA(cudaGetLastError());
    }
    return BLP_OK;
}

// ---- one launch per eval batch ---------------------------------------------------------------------
// Workspace of the fused step: a CTA ticket, then two zero-invariant int32 accumulator arrays of out_len entries.
static long long step_ws_bytes(long long out_len) { return 64 + 8 * (out_len > 0 ? out_len : 0); }
constexpr long long kFusedEpilogueMax = 16384;    // the last CTA walks the whole output range: keep that short

struct StepOut {
    KValues kv; float *recip; unsigned char *hits; double *sums;     // optional metrics (sums == NULL: none)
    void *workspace;
};

// True scores, sweep and (optionally) the metrics of one batch.  d == 128 on a non-empty shard with a workspace:
// ONE launch (true scores per group inside the sweep kernel, last-CTA epilogue); otherwise the separate kernels.
static int rank_step_impl(RankJob j, const StepOut &o, cudaStream_t st) {
    const long long out_len = j.roles == 3 ? j.tail_off + j.b : j.b;  // single-role passes write [0, b) only
    const bool dense_out = j.roles == 3 && j.tail_off == j.b;        // metrics kernels walk [0, 2b)
    const bool tiled = j.d == kD && j.n_local > 0 && aligned16(j.ent) && job_rows_aligned(j) && !j.fast_table_ws;
    if (tiled) {
        SweepArgs a{};
        fill_sweep_args(a, j);
        a.true_score = nullptr;
        a.fuse_true = 1;
        a.true_score_out = j.true_score;
        a.out_len = out_len;
        const bool epi = o.workspace && out_len <= kFusedEpilogueMax && (dense_out || !o.sums);
        if (epi) {
            unsigned char *ws = reinterpret_cast<unsigned char *>(o.workspace);
            a.fuse_epilogue = 1;
            a.overlap = j.overlap;
            a.ticket = reinterpret_cast<unsigned int *>(ws);
            a.gt = reinterpret_cast<int *>(ws + 64);
            a.ge = a.gt + out_len;
            a.gt_out = j.gt; a.ge_out = j.ge;
            a.kv = o.kv; a.recip = o.recip; a.hits = o.hits; a.sums = o.sums;
            return launch_sweep_dyn(j.model, a, st);
        }
        // long output ranges: counters zeroed by a memset, metrics by their own kernel
        if (dense_out || j.roles != 3) {
            BLP_CUDA(cudaMemsetAsync(j.gt, 0, sizeof(int) * out_len, st));
            BLP_CUDA(cudaMemsetAsync(j.ge, 0, sizeof(int) * out_len, st));
        } else {
            BLP_CUDA(cudaMemsetAsync(j.gt, 0, sizeof(int) * j.b, st));
            BLP_CUDA(cudaMemsetAsync(j.gt + j.tail_off, 0, sizeof(int) * j.b, st));
            BLP_CUDA(cudaMemsetAsync(j.ge, 0, sizeof(int) * j.b, st));
            BLP_CUDA(cudaMemsetAsync(j.ge + j.tail_off, 0, sizeof(int) * j.b, st));
        }
        const int rc = launch_sweep_dyn(j.model, a, st);
        if (rc) return rc;
    } else {
        if (j.roles != 3) { set_error("single-role ranking needs d == 128 rows that are 16-byte aligned"); return BLP_EDIM; }
        j.phases = 3;
        const int rc = rank_impl(j, st);
        if (rc) return rc;
    }
    if (o.sums) {
        if (!dense_out) { set_error("metrics need contiguous outputs (tail_off == t)"); return BLP_EINVAL; }
        launch_metrics_reduce(j.gt, j.ge, 2 * j.b, o.kv, o.recip, o.hits, o.sums, st);
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

// gather views of the query rows (train.py:141-143 folded into the kernels)
struct TripleRows { RowRef h, t, r; };
static TripleRows triple_rows(const float *ent, int64_t n_local, int64_t ent_offset, const float *rel_weight, int64_t num_rel,
                              const int64_t *triples, const float *h_rows, const float *t_rows) {
    const long long *tr = (const long long *)triples;
    TripleRows q;
    q.h = h_rows ? dense_rows(h_rows) : RowRef{ent, tr + 0, 3, ent_offset, n_local};
    q.t = t_rows ? dense_rows(t_rows) : RowRef{ent, tr + 1, 3, ent_offset, n_local};
    q.r = RowRef{rel_weight, tr + 2, 3, 0, num_rel};
    return q;
}

static int check_sweep_args(const float *rel_weight, int64_t num_rel, const int64_t *triples, int64_t t, const float *h_rows,
                            const float *t_rows, int64_t tail_off, int64_t n_local) {
    if (!rel_weight || !triples || num_rel <= 0) { set_error("null pointer argument"); return BLP_EINVAL; }
    if ((h_rows == nullptr) != (t_rows == nullptr)) { set_error("h_rows and t_rows must both be given or both NULL"); return BLP_EINVAL; }
    if (tail_off < t) { set_error("tail_off must be >= t"); return BLP_EINVAL; }
    if (!h_rows && n_local == 0) { set_error("cannot gather query rows from an empty shard; pass h_rows / t_rows"); return BLP_EINVAL; }
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
    RankJob j = make_job(model, ent, n_local, ent_offset, d, dense_rows(h_rows), dense_rows(t_rows), dense_rows(r_rows), b, b,
                         gt, ge, true_score);
    j.filt_indptr = (const long long *)filt_indptr; j.filt_idx = (const long long *)filt_idx; j.gt_f = gt_f; j.ge_f = ge_f;
    return rank_impl(j, (cudaStream_t)stream);
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
    if ((rc = check_sweep_args(rel_weight, num_rel, triples, t, h_rows, t_rows, tail_off, n_local))) return rc;
    const TripleRows q = triple_rows(ent, n_local, ent_offset, rel_weight, num_rel, triples, h_rows, t_rows);
    RankJob j = make_job(model, ent, n_local, ent_offset, d, q.h, q.t, q.r, t, tail_off, gt, ge, true_score);
    j.filt_indptr = (const long long *)filt_indptr; j.filt_idx = (const long long *)filt_idx; j.gt_f = gt_f; j.ge_f = ge_f;
    return rank_impl(j, (cudaStream_t)stream);
}

// Chunked sweeps (the reference's eval batches): the true scores and the counter reset of ALL t triples in one
// launch, then one sweep launch per chunk (blp_rank_sweep_counts).
extern "C" int blp_true_scores(int model, const float *ent, int64_t n_local, int64_t ent_offset, int d,
                               const float *rel_weight, int64_t num_rel, const int64_t *triples, int64_t t,
                               const float *h_rows, const float *t_rows, int64_t tail_off, int32_t *gt, int32_t *ge,
                               float *true_score, void *stream) {
    reset_launch_count();
    int rc = check_rank_args(model, d, t, n_local, ent, nullptr, nullptr, gt, ge, nullptr, nullptr, true_score);
    if (rc) return rc;
    if (t == 0) return BLP_OK;
    if ((rc = check_sweep_args(rel_weight, num_rel, triples, t, h_rows, t_rows, tail_off, n_local))) return rc;
    const TripleRows q = triple_rows(ent, n_local, ent_offset, rel_weight, num_rel, triples, h_rows, t_rows);
    RankJob j = make_job(model, ent, n_local, ent_offset, d, q.h, q.t, q.r, t, tail_off, gt, ge, true_score);
    j.phases = 1;
    return rank_impl(j, (cudaStream_t)stream);
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
    if ((rc = check_sweep_args(rel_weight, num_rel, triples, t, h_rows, t_rows, tail_off, n_local))) return rc;
    const TripleRows q = triple_rows(ent, n_local, ent_offset, rel_weight, num_rel, triples, h_rows, t_rows);
    RankJob j = make_job(model, ent, n_local, ent_offset, d, q.h, q.t, q.r, t, tail_off, gt, ge, const_cast<float *>(true_score));
    j.filt_indptr = (const long long *)filt_indptr; j.filt_idx = (const long long *)filt_idx; j.gt_f = gt_f; j.ge_f = ge_f;
    j.phases = 2;
    return rank_impl(j, (cudaStream_t)stream);
}

// ---- the fused step: one launch per eval batch ------------------------------------------------------
static int fill_kvalues(KValues &kv, const int64_t *k_values_host, int nk) {
    if (nk < 0 || nk > 8 || (nk > 0 && !k_values_host)) { set_error("bad k_values (nk <= 8)"); return BLP_EINVAL; }
    kv.nk = nk;
    for (int i = 0; i < nk; ++i) kv.k[i] = k_values_host[i];
    return BLP_OK;
}

extern "C" int64_t blp_rank_step_workspace_bytes(int64_t out_len) { return step_ws_bytes(out_len); }

extern "C" int blp_rank_step(int model, const float *ent, int64_t n_local, int64_t ent_offset, int d,
                             const float *rel_weight, int64_t num_rel, const int64_t *triples, int64_t t,
                             const float *h_rows, const float *t_rows, int64_t tail_off, int group_triples,
                             int32_t *gt, int32_t *ge, float *true_score, const int64_t *k_values_host, int nk,
                             float *recip, uint8_t *hits, double *sums, void *workspace, void *stream) {
    reset_launch_count();
    int rc = check_rank_args(model, d, t, n_local, ent, nullptr, nullptr, gt, ge, nullptr, nullptr, true_score);
    if (rc) return rc;
    if (t == 0) return BLP_OK;
    if ((rc = check_sweep_args(rel_weight, num_rel, triples, t, h_rows, t_rows, tail_off, n_local))) return rc;
    StepOut o{};
    if ((rc = fill_kvalues(o.kv, k_values_host, nk))) return rc;
    o.recip = recip; o.hits = hits; o.sums = sums; o.workspace = workspace;
    const TripleRows q = triple_rows(ent, n_local, ent_offset, rel_weight, num_rel, triples, h_rows, t_rows);
    RankJob j = make_job(model, ent, n_local, ent_offset, d, q.h, q.t, q.r, t, tail_off, gt, ge, true_score);
    j.force_cfg = sweep_cfg_for_group(group_triples);
    return rank_step_impl(j, o, (cudaStream_t)stream);
}

// Everything of blp_rank_step that does not change from batch to batch, bound once (the reference's eval loop calls
// the step every 64 triples: argument marshalling is a visible share of a 20 us step).
struct RankPlan {
    int model; const float *ent; int64_t n_local, ent_offset; int d; const float *rel_weight; int64_t num_rel, t, tail_off;
    int group_triples; int32_t *gt, *ge; float *true_score; int64_t k_values[8]; int nk; float *recip; uint8_t *hits;
    double *sums; void *workspace;
    int overlap;
};

extern "C" int blp_plan_create(void **plan_out, int model, const float *ent, int64_t n_local, int64_t ent_offset, int d,
                               const float *rel_weight, int64_t num_rel, int64_t t, int64_t tail_off, int group_triples,
                               int32_t *gt, int32_t *ge, float *true_score, const int64_t *k_values_host, int nk,
                               float *recip, uint8_t *hits, double *sums, void *workspace) {
    if (!plan_out) { set_error("null plan_out"); return BLP_EINVAL; }
    *plan_out = nullptr;
    int rc = check_rank_args(model, d, t, n_local, ent, nullptr, nullptr, gt, ge, nullptr, nullptr, true_score);
    if (rc) return rc;
    if (nk < 0 || nk > 8 || (nk > 0 && !k_values_host)) { set_error("bad k_values (nk <= 8)"); return BLP_EINVAL; }
    if (!rel_weight || num_rel <= 0 || tail_off < t) { set_error("bad argument"); return BLP_EINVAL; }
    RankPlan *p = new RankPlan{model, ent, n_local, ent_offset, d, rel_weight, num_rel, t, tail_off, group_triples, gt, ge,
                               true_score, {0}, nk, recip, hits, sums, workspace, 0};
    for (int i = 0; i < nk; ++i) p->k_values[i] = k_values_host[i];
    *plan_out = p;
    return BLP_OK;
}

extern "C" int blp_plan_set_overlap(void *plan, int on) {
    if (!plan) { set_error("null plan"); return BLP_EINVAL; }
    reinterpret_cast<RankPlan *>(plan)->overlap = on ? 1 : 0;
    return BLP_OK;
}

extern "C" int blp_plan_run(void *plan, const int64_t *triples, const float *h_rows, const float *t_rows, void *stream) {
    if (!plan) { set_error("null plan"); return BLP_EINVAL; }
    const RankPlan *p = reinterpret_cast<const RankPlan *>(plan);
    if (!p->overlap)
        return blp_rank_step(p->model, p->ent, p->n_local, p->ent_offset, p->d, p->rel_weight, p->num_rel, triples, p->t, h_rows,
                             t_rows, p->tail_off, p->group_triples, p->gt, p->ge, p->true_score, p->k_values, p->nk, p->recip,
                             p->hits, p->sums, p->workspace, stream);
    // as blp_rank_step, launched with programmatic stream serialization
    reset_launch_count();
    int rc = check_rank_args(p->model, p->d, p->t, p->n_local, p->ent, nullptr, nullptr, p->gt, p->ge, nullptr, nullptr, p->true_score);
    if (rc) return rc;
    if (p->t == 0) return BLP_OK;
    if ((rc = check_sweep_args(p->rel_weight, p->num_rel, triples, p->t, h_rows, t_rows, p->tail_off, p->n_local))) return rc;
    StepOut o{};
    if ((rc = fill_kvalues(o.kv, p->k_values, p->nk))) return rc;
    o.recip = p->recip; o.hits = p->hits; o.sums = p->sums; o.workspace = p->workspace;
    const TripleRows q = triple_rows(p->ent, p->n_local, p->ent_offset, p->rel_weight, p->num_rel, triples, h_rows, t_rows);
    RankJob j = make_job(p->model, p->ent, p->n_local, p->ent_offset, p->d, q.h, q.t, q.r, p->t, p->tail_off, p->gt, p->ge, p->true_score);
    j.force_cfg = sweep_cfg_for_group(p->group_triples);
    j.overlap = 1;
    return rank_step_impl(j, o, (cudaStream_t)stream);
}

extern "C" void blp_plan_destroy(void *plan) { delete reinterpret_cast<RankPlan *>(plan); }

// Ranking of unrelated queries against the table (what `get_metrics(score_fn(...), true_idx, k)` asks for when the
// score matrix is never materialised): n_hq head-prediction queries -- candidate row e scored as
// score_fn(e, hq_tails[i], hq_rels[i]), true candidate hq_true[i] -- then n_tq tail-prediction queries
// score_fn(tq_heads[i], e, tq_rels[i]) with true candidate tq_true[i].  Outputs hold the head queries first.
extern "C" int blp_rank_queries(int model, const float *ent, int64_t n_local, int64_t ent_offset, int d,
                                const float *hq_tails, const float *hq_rels, const int64_t *hq_true, int64_t n_hq,
                                const float *tq_heads, const float *tq_rels, const int64_t *tq_true, int64_t n_tq,
                                int32_t *gt, int32_t *ge, float *true_score, const int64_t *k_values_host, int nk,
                                float *recip, uint8_t *hits, double *sums, void *workspace, void *stream) {
    reset_launch_count();
    const int64_t nq = n_hq + n_tq;
    if (n_hq < 0 || n_tq < 0) { set_error("negative size"); return BLP_EINVAL; }
    int rc = check_rank_args(model, d, nq, n_local, ent, nullptr, nullptr, gt, ge, nullptr, nullptr, true_score);
    if (rc) return rc;
    if (nq == 0) return BLP_OK;
    if (n_local <= 0) { set_error("blp_rank_queries gathers the true rows from the table: empty shard"); return BLP_EINVAL; }
    if ((n_hq > 0 && (!hq_tails || !hq_rels || !hq_true)) || (n_tq > 0 && (!tq_heads || !tq_rels || !tq_true))) {
        set_error("null pointer argument");
        return BLP_EINVAL;
    }
    StepOut o{};
    if ((rc = fill_kvalues(o.kv, k_values_host, nk))) return rc;
    o.recip = recip; o.hits = hits; o.sums = sums; o.workspace = workspace;
    cudaStream_t st = (cudaStream_t)stream;
    const RowRef hq_h = RowRef{ent, (const long long *)hq_true, 1, ent_offset, n_local};
    const RowRef tq_t = RowRef{ent, (const long long *)tq_true, 1, ent_offset, n_local};
    if (n_hq == n_tq) {
        RankJob j = make_job(model, ent, n_local, ent_offset, d, hq_h, dense_rows(hq_tails), dense_rows(hq_rels), n_hq, n_hq,
                             gt, ge, true_score);
        j.h2 = dense_rows(tq_heads); j.t2 = tq_t; j.r2 = dense_rows(tq_rels); j.split = 1;
        return rank_step_impl(j, o, st);
    }
    // unequal parts: one single-role pass each, metrics over the concatenation afterwards
    StepOut part = o;
    part.sums = nullptr; part.recip = nullptr; part.hits = nullptr;
    if (n_hq > 0) {
        RankJob j = make_job(model, ent, n_local, ent_offset, d, hq_h, dense_rows(hq_tails), dense_rows(hq_rels), n_hq, n_hq,
                             gt, ge, true_score);
        j.roles = 1;
        if ((rc = rank_step_impl(j, part, st))) return rc;
    }
    if (n_tq > 0) {
        RankJob j = make_job(model, ent, n_local, ent_offset, d, dense_rows(tq_heads), tq_t, dense_rows(tq_rels), n_tq, n_tq,
                             gt + n_hq, ge + n_hq, true_score + n_hq);
        j.roles = 2;
        if ((rc = rank_step_impl(j, part, st))) return rc;
    }
    if (sums) {
        launch_metrics_reduce(gt, ge, nq, o.kv, recip, hits, sums, st);
        count_launch();
        BLP_CUDA(cudaGetLastError());
    }
    return BLP_OK;
}

namespace blp { void set_debug_timestamp_buffer(unsigned long long *p); }
extern "C" int blp_debug_timestamps(void *buffer) {
    blp::set_debug_timestamp_buffer(reinterpret_cast<unsigned long long *>(buffer));
    return BLP_OK;
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

extern "C" int64_t blp_fast_refine_bytes(int64_t capacity) { return fast_refine_ws_bytes(capacity); }

static int rank_sweep_fast_impl(int model, const float *ent, int64_t n_local, int64_t ent_offset, int d,
                                const float *rel_weight, int64_t num_rel, const int64_t *triples, int64_t t,
                                const float *h_rows, const float *t_rows, const int64_t *filt_indptr,
                                const int64_t *filt_idx, int64_t tail_off, int32_t *gt, int32_t *ge, int32_t *gt_f,
                                int32_t *ge_f, float *true_score, const void *table_ws, void *query_ws,
                                float *scores_out, int64_t ld_scores, void *refine_ws, int64_t refine_capacity, void *stream);

extern "C" int blp_rank_sweep_fast(int model, const float *ent, int64_t n_local, int64_t ent_offset, int d,
                                   const float *rel_weight, int64_t num_rel, const int64_t *triples, int64_t t,
                                   const float *h_rows, const float *t_rows, const int64_t *filt_indptr,
                                   const int64_t *filt_idx, int64_t tail_off, int32_t *gt, int32_t *ge, int32_t *gt_f,
                                   int32_t *ge_f, float *true_score, const void *table_ws, void *query_ws,
                                   float *scores_out, int64_t ld_scores, void *stream) {
    return rank_sweep_fast_impl(model, ent, n_local, ent_offset, d, rel_weight, num_rel, triples, t, h_rows, t_rows, filt_indptr,
                                filt_idx, tail_off, gt, ge, gt_f, ge_f, true_score, table_ws, query_ws, scores_out, ld_scores,
                                nullptr, 0, stream);
}

extern "C" int blp_rank_sweep_fast_exact(int model, const float *ent, int64_t n_local, int64_t ent_offset, int d,
                                         const float *rel_weight, int64_t num_rel, const int64_t *triples, int64_t t,
                                         const float *h_rows, const float *t_rows, const int64_t *filt_indptr,
                                         const int64_t *filt_idx, int64_t tail_off, int32_t *gt, int32_t *ge, int32_t *gt_f,
                                         int32_t *ge_f, float *true_score, const void *table_ws, void *query_ws,
                                         void *refine_ws, int64_t refine_capacity, void *stream) {
    if (!refine_ws || refine_capacity <= 0) { set_error("blp_rank_sweep_fast_exact needs a refine workspace"); return BLP_EINVAL; }
    if (refine_capacity >= (1ll << 31)) { set_error("refine_capacity must be < 2^31"); return BLP_EINVAL; }
    if (n_local >= (1ll << 31) || 2 * t >= (1ll << 31)) { set_error("refine entries are int32 (query, candidate) pairs"); return BLP_EINVAL; }
    if (!aligned16(ent) || !aligned16(rel_weight) || !aligned16(h_rows) || !aligned16(t_rows) || !aligned16(refine_ws)) {
        set_error("ent / rel_weight / h_rows / t_rows / refine_ws must be 16-byte aligned");
        return BLP_EINVAL;
    }
    return rank_sweep_fast_impl(model, ent, n_local, ent_offset, d, rel_weight, num_rel, triples, t, h_rows, t_rows, filt_indptr,
                                filt_idx, tail_off, gt, ge, gt_f, ge_f, true_score, table_ws, query_ws, nullptr, 0, refine_ws,
                                refine_capacity, stream);
}

static int rank_sweep_fast_impl(int model, const float *ent, int64_t n_local, int64_t ent_offset, int d,
                                const float *rel_weight, int64_t num_rel, const int64_t *triples, int64_t t,
                                const float *h_rows, const float *t_rows, const int64_t *filt_indptr,
                                const int64_t *filt_idx, int64_t tail_off, int32_t *gt, int32_t *ge, int32_t *gt_f,
                                int32_t *ge_f, float *true_score, const void *table_ws, void *query_ws,
                                float *scores_out, int64_t ld_scores, void *refine_ws, int64_t refine_capacity, void *stream) {
    reset_launch_count();
    int rc = check_rank_args(model, d, t, n_local, ent, filt_indptr, filt_idx, gt, ge, gt_f, ge_f, true_score);
    if (rc) return rc;
    if (d != kD) { set_error("fast mode is specialised for d = %d (got %d)", kD, d); return BLP_EDIM; }
    if (model == BLP_MODEL_TRANSE) { set_error("fast (tensor-core) mode covers distmult / complex / simple only"); return BLP_EINVAL; }
    if (t == 0) return BLP_OK;
    if (!table_ws || !query_ws) { set_error("null pointer argument"); return BLP_EINVAL; }
    if ((rc = check_sweep_args(rel_weight, num_rel, triples, t, h_rows, t_rows, tail_off, n_local))) return rc;
    if (!aligned16(table_ws) || !aligned16(query_ws)) { set_error("workspaces must be 16-byte aligned"); return BLP_EINVAL; }
    if (scores_out && ld_scores < n_local) { set_error("ld_scores must be >= n_local"); return BLP_EINVAL; }
    const TripleRows q = triple_rows(ent, n_local, ent_offset, rel_weight, num_rel, triples, h_rows, t_rows);
    RankJob j = make_job(model, ent, n_local, ent_offset, d, q.h, q.t, q.r, t, tail_off, gt, ge, true_score);
    j.filt_indptr = (const long long *)filt_indptr; j.filt_idx = (const long long *)filt_idx; j.gt_f = gt_f; j.ge_f = ge_f;
    j.triples = (const long long *)triples; j.fast_table_ws = table_ws; j.fast_query_ws = query_ws; j.fast_scores = scores_out;
    j.fast_ld = ld_scores;
    j.fast_refine_ws = refine_ws; j.fast_refine_cap = refine_capacity;
    return rank_impl(j, (cudaStream_t)stream);
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
        a.h2 = a.h; a.t2 = a.t; a.r2 = a.r; a.split_sets = 0; a.force_cfg = -1;
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
    launch_metrics_reduce(gt, ge, q, kv, recip, hits, sums, (cudaStream_t)stream);
    count_launch();
    BLP_CUDA(cudaGetLastError());
    return BLP_OK;
}
