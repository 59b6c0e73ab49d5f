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


struct SweepArgs {
    const float *ent;        // [n_local, 128]
    long long n_local;
    RowRef h, t, r;          // true head / true tail / relation row of triple i
    long long b;
    long long tail_off;      // outputs: head query i -> [i], tail query i -> [tail_off + i]
    const float *true_score; // indexed like the outputs, or NULL when writing scores
    int *gt;
    int *ge;
    float *scores_out;       // optional (n_queries, ld_scores) matrix instead of counting
    long long ld_scores;
    int roles;               // 3 = both, 1 = head prediction only, 2 = tail prediction only
    long long groups;        // triple groups, filled in by the launcher
    int use_tma;             // natural-order tiles through TMA bulk copies (TransE)
    unsigned long long negzero2;   // kNegZero2, opaque to ptxas (see mul2)
};

int launch_sweep_dyn(int model, const SweepArgs &a, cudaStream_t st);

// tensor-core fast mode (blp_fast.cu)
long long fast_table_ws_bytes(long long n_local);
long long fast_query_ws_bytes(long long t);
int fast_prepare_table(const float *ent, long long n_local, void *table_ws, cudaStream_t st);
int launch_fast_sweep(int model, long long n_local, long long ent_offset, const RowRef &h, const RowRef &t, const RowRef &r,
                      const long long *triples, long long b, long long tail_off, float *true_score, int *gt, int *ge,
                      const void *table_ws, void *query_ws, float *scores_out, long long ld_scores, bool compute_true,
                      cudaStream_t st);
int sweep_env_use_tma();
constexpr long long kSweepMaxB = 1ll << 40;

}  // namespace blp
