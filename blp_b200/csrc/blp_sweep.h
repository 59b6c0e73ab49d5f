// blp_sweep.h -- interface between the C-ABI entry points (blp_eval.cu) and the d == 128 sweep kernel (blp_sweep.cu).
#pragma once
#include "blp_common.cuh"

namespace blp {

constexpr int kD = 128;          // specialised row width

// Row i of a query operand: either the i-th row of a dense [b, d] block (idx == NULL) or a gather
// base[idx[i * idx_stride] - offset] (train.py:141-143 folded into the kernels).  Out-of-range indices are
// clamped (the true-score kernel flags them with a NaN score).
struct RowRef {
    const float *base;
    const long long *idx;
    long long idx_stride, offset, limit;
    __host__ __device__ __forceinline__ bool in_range(long long i) const {
        if (!idx) return true;
        const long long r = idx[i * idx_stride] - offset;
        return r >= 0 && r < limit;
    }
    __host__ __device__ __forceinline__ const float *row(long long i, int d) const {
        if (!idx) return base + i * d;
        long long r = idx[i * idx_stride] - offset;
        r = r < 0 ? 0 : (r >= limit ? limit - 1 : r);
        return base + r * d;
    }
};
inline RowRef dense_rows(const float *p) { return RowRef{p, nullptr, 0, 0, 0}; }

// Exact score of one triple at d == 128 by one warp (SURVEY.md Appendix A bits, ~10x shorter critical path than the
// single-thread score_exact): the per-position terms are computed in parallel, one float4 (float2 for the halves
// models) per lane, and parked in `tm` (128 floats of shared memory private to the warp); only the reference's
// summation order is replayed serially -- 128 dependent adds for torch.norm(p=1) (lane 0), or ATen's 8-lane x
// 4-accumulator cascade for torch.sum (lanes 0-7 run their chains in parallel, the 8 lane sums fold sequentially).
// The result is valid in lane 0 (all lanes for the bilinear models).  h, t, r must be 16-byte aligned.
template <int MODEL>
__device__ __forceinline__ float true_score_warp128(const float *__restrict__ h, const float *__restrict__ t,
                                                    const float *__restrict__ r, float *__restrict__ tm, int lane) {
    constexpr int L = (MODEL == BLP_MODEL_COMPLEX || MODEL == BLP_MODEL_SIMPLE) ? kD / 2 : kD;
    if (MODEL == BLP_MODEL_TRANSE || MODEL == BLP_MODEL_DISTMULT) {
        const float4 hv = __ldg(reinterpret_cast<const float4 *>(h) + lane), tv = __ldg(reinterpret_cast<const float4 *>(t) + lane),
                     rv = __ldg(reinterpret_cast<const float4 *>(r) + lane);
        float4 o;
        if (MODEL == BLP_MODEL_TRANSE) {
            o.x = fabsf(fsub(fadd(hv.x, rv.x), tv.x)); o.y = fabsf(fsub(fadd(hv.y, rv.y), tv.y));
            o.z = fabsf(fsub(fadd(hv.z, rv.z), tv.z)); o.w = fabsf(fsub(fadd(hv.w, rv.w), tv.w));
        } else {
            o.x = fmul(fmul(hv.x, rv.x), tv.x); o.y = fmul(fmul(hv.y, rv.y), tv.y);
            o.z = fmul(fmul(hv.z, rv.z), tv.z); o.w = fmul(fmul(hv.w, rv.w), tv.w);
        }
        reinterpret_cast<float4 *>(tm)[lane] = o;
    } else {
        // halves: lane owns positions 2 * lane, 2 * lane + 1 of [0, 64)
        const float2 h0 = __ldg(reinterpret_cast<const float2 *>(h) + lane), h1 = __ldg(reinterpret_cast<const float2 *>(h + L) + lane);
        const float2 t0 = __ldg(reinterpret_cast<const float2 *>(t) + lane), t1 = __ldg(reinterpret_cast<const float2 *>(t + L) + lane);
        const float2 r0 = __ldg(reinterpret_cast<const float2 *>(r) + lane), r1 = __ldg(reinterpret_cast<const float2 *>(r + L) + lane);
        const float hx[2] = {h0.x, h0.y}, hy[2] = {h1.x, h1.y}, tx[2] = {t0.x, t0.y}, ty[2] = {t1.x, t1.y};
        const float rx[2] = {r0.x, r0.y}, ry[2] = {r1.x, r1.y};
        float o[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (MODEL == BLP_MODEL_COMPLEX) {        // models.py:230-239, left to right
                float p = fadd(fmul(fmul(rx[u], hx[u]), tx[u]), fmul(fmul(rx[u], hy[u]), ty[u]));
                p = fadd(p, fmul(fmul(ry[u], hx[u]), ty[u]));
                o[u] = fsub(p, fmul(fmul(ry[u], hy[u]), tx[u]));
            } else {                                  // models.py:242-248
                o[u] = fadd(fmul(fmul(hx[u], rx[u]), ty[u]), fmul(fmul(tx[u], ry[u]), hy[u]));
            }
        }
        reinterpret_cast<float2 *>(tm)[lane] = make_float2(o[0], o[1]);
    }
    __syncwarp();
    float s = 0.0f;
    if (MODEL == BLP_MODEL_TRANSE) {
        if (lane == 0) {
#pragma unroll 8
            for (int j = 0; j < kD; j += 4) {
                const float4 v = *reinterpret_cast<const float4 *>(tm + j);
                s = fadd(fadd(fadd(fadd(s, v.x), v.y), v.z), v.w);
            }
            s = -s;
        }
    } else {
        // element j -> lane j % 8, accumulator (j / 8) % 4, in increasing j (L / 32 steps per accumulator)
        float c = 0.0f;
        if (lane < 8) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < L / 32; ++k)
#pragma unroll
                for (int a = 0; a < 4; ++a) acc[a] = fadd(acc[a], tm[32 * k + 8 * a + lane]);
            c = fadd(fadd(fadd(acc[0], acc[1]), acc[2]), acc[3]);
        }
#pragma unroll
        for (int l = 0; l < 8; ++l) s = fadd(s, __shfl_sync(0xffffffffu, c, l));
        if (MODEL == BLP_MODEL_SIMPLE) s = fmul(s, 0.5f);
    }
    return s;
}


// Up to 4 exact true-triple scores at d == 128 by one warp, chains run side by side: job u (u < n) scores
// (h[u], t[u], r[u]).  Same operations in the same order as true_score_warp128, but the serial
// part -- 128 dependent adds for torch.norm(p=1), ATen's 8-lane x 4-accumulator cascade for torch.sum -- of the
// jobs runs in different lanes at the same time (lane u for TransE, lanes 8u .. 8u+7 for the bilinear models).
// `tm` = 4 x 128 floats of shared memory private to the warp.  On return lane u (TransE) / every lane of
// 8u .. 8u+7 (bilinear) holds the score of job u; use job_lane() to read it.
template <int MODEL>
__device__ __forceinline__ int true_job_lane(int u) { return MODEL == BLP_MODEL_TRANSE ? u : 8 * u; }

template <int MODEL>
__device__ __forceinline__ float true_scores_warp128x4(const float *const (&h)[4], const float *const (&t)[4],
                                                       const float *const (&r)[4], int n, float *__restrict__ tm, int lane) {
    constexpr int L = (MODEL == BLP_MODEL_COMPLEX || MODEL == BLP_MODEL_SIMPLE) ? kD / 2 : kD;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        if (u >= n) break;
        float *tu = tm + u * kD;
        if (MODEL == BLP_MODEL_TRANSE || MODEL == BLP_MODEL_DISTMULT) {
            const float4 hv = __ldg(reinterpret_cast<const float4 *>(h[u]) + lane), tv = __ldg(reinterpret_cast<const float4 *>(t[u]) + lane),
                         rv = __ldg(reinterpret_cast<const float4 *>(r[u]) + lane);
            float4 o;
            if (MODEL == BLP_MODEL_TRANSE) {
                o.x = fabsf(fsub(fadd(hv.x, rv.x), tv.x)); o.y = fabsf(fsub(fadd(hv.y, rv.y), tv.y));
                o.z = fabsf(fsub(fadd(hv.z, rv.z), tv.z)); o.w = fabsf(fsub(fadd(hv.w, rv.w), tv.w));
            } else {
                o.x = fmul(fmul(hv.x, rv.x), tv.x); o.y = fmul(fmul(hv.y, rv.y), tv.y);
                o.z = fmul(fmul(hv.z, rv.z), tv.z); o.w = fmul(fmul(hv.w, rv.w), tv.w);
            }
            reinterpret_cast<float4 *>(tu)[lane] = o;
        } else {
            const float2 h0 = __ldg(reinterpret_cast<const float2 *>(h[u]) + lane), h1 = __ldg(reinterpret_cast<const float2 *>(h[u] + L) + lane);
            const float2 t0 = __ldg(reinterpret_cast<const float2 *>(t[u]) + lane), t1 = __ldg(reinterpret_cast<const float2 *>(t[u] + L) + lane);
            const float2 r0 = __ldg(reinterpret_cast<const float2 *>(r[u]) + lane), r1 = __ldg(reinterpret_cast<const float2 *>(r[u] + L) + lane);
            const float hx[2] = {h0.x, h0.y}, hy[2] = {h1.x, h1.y}, tx[2] = {t0.x, t0.y}, ty[2] = {t1.x, t1.y};
            const float rx[2] = {r0.x, r0.y}, ry[2] = {r1.x, r1.y};
            float o[2];
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                if (MODEL == BLP_MODEL_COMPLEX) {        // models.py:230-239, left to right
                    float p = fadd(fmul(fmul(rx[v], hx[v]), tx[v]), fmul(fmul(rx[v], hy[v]), ty[v]));
                    p = fadd(p, fmul(fmul(ry[v], hx[v]), ty[v]));
                    o[v] = fsub(p, fmul(fmul(ry[v], hy[v]), tx[v]));
                } else {                                  // models.py:242-248
                    o[v] = fadd(fmul(fmul(hx[v], rx[v]), ty[v]), fmul(fmul(tx[v], ry[v]), hy[v]));
                }
            }
            reinterpret_cast<float2 *>(tu)[lane] = make_float2(o[0], o[1]);
        }
    }
    __syncwarp();
    float s = 0.0f;
    if (MODEL == BLP_MODEL_TRANSE) {
        if (lane < n) {
            const float *tu = tm + lane * kD;
#pragma unroll 8
            for (int j = 0; j < kD; j += 4) {
                const float4 v = *reinterpret_cast<const float4 *>(tu + j);
                s = fadd(fadd(fadd(fadd(s, v.x), v.y), v.z), v.w);
            }
            s = -s;
        }
    } else {
        // lane = 8 * job + l: element j -> lane l = j % 8, accumulator (j / 8) % 4, in increasing j
        const int u = lane >> 3, l = lane & 7;
        const float *tu = tm + u * kD;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (u < n) {
#pragma unroll
            for (int k = 0; k < L / 32; ++k)
#pragma unroll
                for (int a = 0; a < 4; ++a) acc[a] = fadd(acc[a], tu[32 * k + 8 * a + l]);
        }
        const float c = fadd(fadd(fadd(acc[0], acc[1]), acc[2]), acc[3]);
#pragma unroll
        for (int ll = 0; ll < 8; ++ll) s = fadd(s, __shfl_sync(0xffffffffu, c, (lane & 24) + ll));
        if (MODEL == BLP_MODEL_SIMPLE) s = fmul(s, 0.5f);
    }
    __syncwarp();
    return s;
}

struct KValues { long long k[8]; int nk; };

struct SweepArgs {
    const float *ent;        // [n_local, 128]
    long long n_local;
    RowRef h, t, r;          // true head / true tail / relation row of triple i (head-prediction queries)
    RowRef h2, t2, r2;       // the same for the tail-prediction queries (== h, t, r when split_sets == 0)
    int split_sets;          // 1: head query i and tail query i are unrelated queries (generic score-matrix path)
    long long b;
    long long tail_off;      // outputs: head query i -> [i], tail query i -> [tail_off + i]
    const float *true_score; // indexed like the outputs; NULL when writing scores or when fuse_true is set
    int *gt;                 // counters the kernel adds into (zero on entry): the caller's arrays, or the
    int *ge;                 //   zero-invariant workspace accumulators when the fused epilogue runs
    float *scores_out;       // optional (n_queries, ld_scores) matrix instead of counting
    long long ld_scores;
    int roles;               // 3 = both, 1 = head prediction only, 2 = tail prediction only
    long long groups;        // triple groups, filled in by the launcher
    int use_tma;             // natural-order tiles through TMA bulk copies (TransE)
    int force_cfg;           // -1 = pick the register tile by batch size; 0..4 = Cfg index (triples per group 2/4/8/16/32)
    unsigned long long negzero2;   // kNegZero2, opaque to ptxas (see mul2)
    // ---- fused step (one launch per eval batch): true scores computed per group inside the kernel, counters
    // accumulated in a zero-invariant workspace, the last CTA to finish writes the results and the metrics
    int fuse_true;           // compute s_true in the kernel (true_score_out receives it)
    float *true_score_out;
    int fuse_epilogue;       // last-CTA epilogue: acc -> gt_out / ge_out (+ recip / hits / sums), acc re-zeroed
    unsigned int *ticket;    // workspace: CTA completion counter (zero on entry, zero on exit)
    int *gt_out, *ge_out;
    KValues kv;
    float *recip;            // optional [.. like outputs], utils.py:106-108
    unsigned char *hits;     // optional [slot * nk + j], utils.py:109
    double *sums;            // optional [1 + nk], train.py:154-157
    long long out_len;       // tail_off + b: length of the output / accumulator index space
    unsigned long long *dbg; // measurement aid (blp_debug_timestamps): 16 globaltimer slots per CTA, or NULL
    int overlap;             // launch with programmatic stream serialization (blp_plan_set_overlap)
};

int launch_sweep_dyn(int model, const SweepArgs &a, cudaStream_t st);
// TransE at any row width d % 4 == 0 (blp_sweep_wide.cu): counting only, true_score given, counters zeroed
bool sweep_wide_supports(int model, int d, const float *ent);
int launch_sweep_wide(const SweepArgs &a, int d, cudaStream_t st);

// tensor-core fast mode (blp_fast.cu)
long long fast_table_ws_bytes(long long n_local);
long long fast_query_ws_bytes(long long t);
long long fast_refine_ws_bytes(long long capacity);
int fast_prepare_table(const float *ent, long long n_local, void *table_ws, cudaStream_t st);
int launch_fast_sweep(int model, long long n_local, long long ent_offset, const RowRef &h, const RowRef &t, const RowRef &r,
                      const long long *triples, long long b, long long tail_off, float *true_score, int *gt, int *ge,
                      const void *table_ws, void *query_ws, float *scores_out, long long ld_scores, bool compute_true,
                      const float *ent, void *refine_ws, long long refine_cap, cudaStream_t st);
int sweep_env_use_tma();
unsigned long long *debug_timestamp_buffer();
int sweep_cfg_for_group(long long group_triples);
constexpr long long kSweepMaxB = 1ll << 40;

}  // namespace blp
