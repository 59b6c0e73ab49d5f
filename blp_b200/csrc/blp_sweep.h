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
                      const long long *triples, long long b, long long tail_off, const float *true_score, int *gt, int *ge,
                      const void *table_ws, void *query_ws, float *scores_out, long long ld_scores, cudaStream_t st);
int sweep_env_use_tma();
constexpr long long kSweepMaxB = 1ll << 40;

}  // namespace blp
