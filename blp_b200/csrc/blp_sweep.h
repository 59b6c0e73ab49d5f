// blp_sweep.h -- interface between the C-ABI entry points (blp_eval.cu) and the d == 128 sweep kernel (blp_sweep.cu).
#pragma once
#include "blp_common.cuh"

namespace blp {

constexpr int kD = 128;          // specialised row width

struct SweepArgs {
    const float *ent;        // [n_local, 128]
    long long n_local;
    const float *h_rows;     // [b, 128]
    const float *t_rows;
    const float *r_rows;
    long long b;
    const float *true_score; // [2b] (head queries then tail queries) or NULL when writing scores
    int *gt;                 // [2b]
    int *ge;
    float *scores_out;       // optional (n_queries, ld_scores) matrix instead of counting
    long long ld_scores;
    int roles;               // 3 = both, 1 = head prediction only, 2 = tail prediction only
    long long groups;        // triple groups, filled in by the launcher
    int use_tma;             // natural-order tiles through TMA bulk copies (TransE)
    unsigned long long negzero2;   // kNegZero2, opaque to ptxas (see mul2)
};

int launch_sweep_dyn(int model, const SweepArgs &a, cudaStream_t st);
int sweep_env_use_tma();
constexpr long long kSweepMaxB = 1ll << 40;

}  // namespace blp
