// blp_eval.cu -- fused full-entity scoring + rank counting (a1-a4, a10-a12).
//
// Replaces train.py:141-171 + utils.py:86-111 of the reference: for every query
// (a test triple with its head or its tail removed) score ALL candidate entity
// rows and reduce the scores to two integers, gt = #{s_j > s_true} and
// ge = #{s_j >= s_true}.  Neither the (B, N, D) broadcast nor the (2B, N) score
// matrix is ever materialised.
//
// Kernel layout (d == 128, the width of every BLP script; other widths take the
// generic kernel at the bottom):
//   - one CTA = 8 consumer warps + 1 producer warp, persistent over candidate
//     tiles; a CTA owns a group of 64 same-role queries (8 per consumer warp);
//   - query rows (h, t, r) are staged into shared memory by TMA bulk copies and
//     folded into <= 2 operand vectors per query (e.g. u = fl(h + r) for TransE
//     tail prediction) stored in *processing order*;
//   - candidate tiles (128 rows) are double-buffered in shared memory with a
//     132-float pitch, so a lane that walks one row with 128-bit loads never
//     conflicts; the producer warp fills them either with TMA bulk copies
//     (natural order: TransE) or with coalesced 128-bit global loads + a
//     permuting store (ATen sum order: DistMult / ComplEx / SimplE);
//   - each consumer thread owns an 8-query x 4-candidate register tile and
//     replays the reference's exact fp32 operation order for each pair
//     (SURVEY.md Appendix A), then compares against the true-triple score.
// The path is FP32-ALU bound for large query groups (2-3 lane-ops per
// (query, candidate, dim)) and HBM bound only for tiny ones; see DESIGN.md.
#include "blp_common.cuh"

namespace blp {

constexpr int kD = 128;                       // specialised row width
constexpr int kPitch = 132;                   // smem row pitch (floats)
constexpr int kTQ = 8;                        // queries per consumer thread
constexpr int kTC = 4;                        // candidates per consumer thread
constexpr int kConsumerWarps = 8;
constexpr int kQT = kTQ * kConsumerWarps;     // 64 queries per CTA
constexpr int kCT = 32 * kTC;                 // 128 candidate rows per tile
constexpr int kThreads = (kConsumerWarps + 1) * 32;
constexpr int kStages = 2;

struct __align__(16) SweepSmem {
    float qv[kQT][2][kD];                     // per-query operand vectors, processing order   (64 KB)
    float ctile[kStages][kCT][kPitch];        // candidate tiles, processing order            (132 KB)
    float st[kQT];                            // true-triple scores of this CTA's queries
    uint64_t full_bar[kStages];
    uint64_t empty_bar[kStages];
    uint64_t q_bar;
};

// position of natural element j inside a staged row
template <int MODEL>
__host__ __device__ __forceinline__ constexpr int perm_pos(int j) {
    if (MODEL == BLP_MODEL_TRANSE) return j;
    if (MODEL == BLP_MODEL_DISTMULT) return (j & 7) * 16 + ((j >> 3) & 3) * 4 + (j >> 5);
    return (j >> 6) * 64 + (j & 7) * 8 + ((j >> 3) & 3) * 2 + ((j & 63) >> 5);
}

template <int MODEL>
struct ModelTraits {
    static constexpr bool kHalves = (MODEL == BLP_MODEL_COMPLEX || MODEL == BLP_MODEL_SIMPLE);
};

// ---- query-side folding -----------------------------------------------------
// HEAD_PRED: the candidate row plays `heads` (train.py:146); otherwise `tails` (train.py:147).
template <int MODEL, bool HEAD_PRED>
__device__ __forceinline__ void fold_query(const float *__restrict__ h, const float *__restrict__ t,
                                           const float *__restrict__ r, int j, float *__restrict__ v0,
                                           float *__restrict__ v1) {
    const int p = perm_pos<MODEL>(j);
    if (MODEL == BLP_MODEL_TRANSE || MODEL == BLP_MODEL_DISTMULT) {
        if (HEAD_PRED) {
            v0[p] = r[j];
            v1[p] = t[j];
        } else {
            v0[p] = (MODEL == BLP_MODEL_TRANSE) ? fadd(h[j], r[j]) : fmul(h[j], r[j]);
            v1[p] = 0.0f;
        }
    } else if (MODEL == BLP_MODEL_COMPLEX) {
        if (HEAD_PRED) {
            v0[p] = r[j];
            v1[p] = t[j];
        } else if (j < 64) {
            const float hr = h[j], hi = h[64 + j], rr = r[j], ri = r[64 + j];
            v0[p] = fmul(rr, hr);        // A
            v0[64 + p] = fmul(rr, hi);   // B
            v1[p] = fmul(ri, hr);        // C
            v1[64 + p] = fmul(ri, hi);   // D
        }
    } else {  // SIMPLE
        if (j < 64) {
            if (HEAD_PRED) {             // candidate = (hh, ht)
                v0[p] = r[j];                          // ra
                v0[64 + p] = t[64 + j];                // tt
                v1[p] = fmul(t[j], r[64 + j]);         // th * rb
                v1[64 + p] = 0.0f;
            } else {                     // candidate = (th, tt)
                v0[p] = fmul(h[j], r[j]);              // hh * ra
                v0[64 + p] = r[64 + j];                // rb
                v1[p] = h[64 + j];                     // ht
                v1[64 + p] = 0.0f;
            }
        }
    }
}

// ---- per-position arithmetic -----------------------------------------------
template <bool HEAD_PRED>
__device__ __forceinline__ float transe_step(float acc, float e, float v0, float v1) {
    const float x = HEAD_PRED ? fsub(fadd(e, v0), v1) : fsub(v0, e);
    return fadd(acc, fabsf(x));
}
template <bool HEAD_PRED>
__device__ __forceinline__ float distmult_term(float e, float v0, float v1) {
    return HEAD_PRED ? fmul(fmul(e, v0), v1) : fmul(v0, e);
}
template <int MODEL, bool HEAD_PRED>
__device__ __forceinline__ float halves_term(float elo, float ehi, float v0lo, float v0hi, float v1lo, float v1hi) {
    if (MODEL == BLP_MODEL_COMPLEX) {
        if (HEAD_PRED) {  // e = (hr, hi); v0 = (rr, ri); v1 = (tr, ti)
            float p = fadd(fmul(fmul(v0lo, elo), v1lo), fmul(fmul(v0lo, ehi), v1hi));
            p = fadd(p, fmul(fmul(v0hi, elo), v1hi));
            return fsub(p, fmul(fmul(v0hi, ehi), v1lo));
        } else {          // e = (tr, ti); v0 = (A, B); v1 = (C, D)
            float p = fadd(fmul(v0lo, elo), fmul(v0hi, ehi));
            p = fadd(p, fmul(v1lo, ehi));
            return fsub(p, fmul(v1hi, elo));
        }
    } else {
        if (HEAD_PRED) return fadd(fmul(fmul(elo, v0lo), v0hi), fmul(v1lo, ehi));   // (hh*ra)*tt + (th*rb)*ht
        return fadd(fmul(v0lo, ehi), fmul(fmul(elo, v0hi), v1lo));                  // (hh*ra)*tt + (th*rb)*ht
    }
}

__device__ __forceinline__ float4 lds128(const float *p) { return *reinterpret_cast<const float4 *>(p); }

// Scores of an 8-query x 4-candidate register tile against one staged tile.
template <int MODEL, bool HEAD_PRED>
__device__ __forceinline__ void score_tile(const float *__restrict__ ct, const float *__restrict__ qv, int lane,
                                           float (&s)[kTQ][kTC]) {
    const float *row0 = ct + lane * kPitch;
#pragma unroll
    for (int q = 0; q < kTQ; ++q)
#pragma unroll
        for (int i = 0; i < kTC; ++i) s[q][i] = 0.0f;

    if (MODEL == BLP_MODEL_TRANSE) {
        // strictly sequential L1 accumulation, natural order
#pragma unroll 2
        for (int c4 = 0; c4 < kD / 4; ++c4) {
            float4 e[kTC];
#pragma unroll
            for (int i = 0; i < kTC; ++i) e[i] = lds128(row0 + i * 32 * kPitch + 4 * c4);
#pragma unroll
            for (int q = 0; q < kTQ; ++q) {
                const float4 a = lds128(qv + q * 2 * kD + 4 * c4);
                float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                if (HEAD_PRED) b = lds128(qv + q * 2 * kD + kD + 4 * c4);
#pragma unroll
                for (int i = 0; i < kTC; ++i) {
                    float acc = s[q][i];
                    acc = transe_step<HEAD_PRED>(acc, e[i].x, a.x, b.x);
                    acc = transe_step<HEAD_PRED>(acc, e[i].y, a.y, b.y);
                    acc = transe_step<HEAD_PRED>(acc, e[i].z, a.z, b.z);
                    acc = transe_step<HEAD_PRED>(acc, e[i].w, a.w, b.w);
                    s[q][i] = acc;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < kTQ; ++q)
#pragma unroll
            for (int i = 0; i < kTC; ++i) s[q][i] = -s[q][i];
    } else if (MODEL == BLP_MODEL_DISTMULT) {
        // staged order: pos = l*16 + a*4 + k  <->  j = 32k + 8a + l; one 16-byte chunk = one (l, a) chain
        for (int l = 0; l < 8; ++l) {
            float c[kTQ][kTC];
#pragma unroll
            for (int q = 0; q < kTQ; ++q)
#pragma unroll
                for (int i = 0; i < kTC; ++i) c[q][i] = 0.0f;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int off = l * 16 + a * 4;
                float4 e[kTC];
#pragma unroll
                for (int i = 0; i < kTC; ++i) e[i] = lds128(row0 + i * 32 * kPitch + off);
#pragma unroll
                for (int q = 0; q < kTQ; ++q) {
                    const float4 v0 = lds128(qv + q * 2 * kD + off);
                    float4 v1 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (HEAD_PRED) v1 = lds128(qv + q * 2 * kD + kD + off);
#pragma unroll
                    for (int i = 0; i < kTC; ++i) {
                        float chain = distmult_term<HEAD_PRED>(e[i].x, v0.x, v1.x);
                        chain = fadd(chain, distmult_term<HEAD_PRED>(e[i].y, v0.y, v1.y));
                        chain = fadd(chain, distmult_term<HEAD_PRED>(e[i].z, v0.z, v1.z));
                        chain = fadd(chain, distmult_term<HEAD_PRED>(e[i].w, v0.w, v1.w));
                        c[q][i] = fadd(c[q][i], chain);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < kTQ; ++q)
#pragma unroll
                for (int i = 0; i < kTC; ++i) s[q][i] = fadd(s[q][i], c[q][i]);
        }
    } else {
        // halves; staged order per half: pos = l*8 + a*2 + k; one chunk = two (l, a) chains of length 2
        for (int l = 0; l < 8; ++l) {
            float c[kTQ][kTC];
#pragma unroll
            for (int q = 0; q < kTQ; ++q)
#pragma unroll
                for (int i = 0; i < kTC; ++i) c[q][i] = 0.0f;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int off = l * 8 + hh * 4;
                float4 elo[kTC], ehi[kTC];
#pragma unroll
                for (int i = 0; i < kTC; ++i) {
                    elo[i] = lds128(row0 + i * 32 * kPitch + off);
                    ehi[i] = lds128(row0 + i * 32 * kPitch + 64 + off);
                }
#pragma unroll
                for (int q = 0; q < kTQ; ++q) {
                    const float4 v0lo = lds128(qv + q * 2 * kD + off);
                    const float4 v0hi = lds128(qv + q * 2 * kD + 64 + off);
                    const float4 v1lo = lds128(qv + q * 2 * kD + kD + off);
                    float4 v1hi = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (MODEL == BLP_MODEL_COMPLEX) v1hi = lds128(qv + q * 2 * kD + kD + 64 + off);
#pragma unroll
                    for (int i = 0; i < kTC; ++i) {
                        const float p0 = halves_term<MODEL, HEAD_PRED>(elo[i].x, ehi[i].x, v0lo.x, v0hi.x, v1lo.x, v1hi.x);
                        const float p1 = halves_term<MODEL, HEAD_PRED>(elo[i].y, ehi[i].y, v0lo.y, v0hi.y, v1lo.y, v1hi.y);
                        const float p2 = halves_term<MODEL, HEAD_PRED>(elo[i].z, ehi[i].z, v0lo.z, v0hi.z, v1lo.z, v1hi.z);
                        const float p3 = halves_term<MODEL, HEAD_PRED>(elo[i].w, ehi[i].w, v0lo.w, v0hi.w, v1lo.w, v1hi.w);
                        float cc = fadd(c[q][i], fadd(p0, p1));
                        c[q][i] = fadd(cc, fadd(p2, p3));
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < kTQ; ++q)
#pragma unroll
                for (int i = 0; i < kTC; ++i) s[q][i] = fadd(s[q][i], c[q][i]);
        }
        if (MODEL == BLP_MODEL_SIMPLE) {
#pragma unroll
            for (int q = 0; q < kTQ; ++q)
#pragma unroll
                for (int i = 0; i < kTC; ++i) s[q][i] = fmul(s[q][i], 0.5f);
        }
    }
}

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
    int groups_per_role;
    int use_tma;             // natural-order tiles through TMA bulk copies (TransE)
};

template <int MODEL>
__global__ void __launch_bounds__(kThreads, 1) sweep_kernel(const SweepArgs args) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SweepSmem &sm = *reinterpret_cast<SweepSmem *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // which query group / role
    int g = blockIdx.y;
    bool head_pred;
    if (args.roles == 3) {
        head_pred = g < args.groups_per_role;
        if (!head_pred) g -= args.groups_per_role;
    } else {
        head_pred = args.roles == 1;
    }
    const long long q0 = (long long)g * kQT;                       // first triple of this group
    const int nq = (int)min((long long)kQT, args.b - q0);
    const long long qout0 = (head_pred || args.roles != 3 ? 0 : args.b) + q0;   // index into gt/ge/true_score

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&sm.full_bar[s], 1);
            mbar_init(&sm.empty_bar[s], kConsumerWarps);
        }
        mbar_init(&sm.q_bar, 1);
        mbar_fence_init();
    }
    __syncthreads();

    // ---- stage raw query rows through TMA into the (still unused) tile buffers
    float *raw_h = &sm.ctile[0][0][0];
    float *raw_t = raw_h + kQT * kD;
    float *raw_r = raw_t + kQT * kD;
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)nq * kD * 4u;
        mbar_arrive_expect_tx(&sm.q_bar, 3u * bytes);
        tma_bulk_g2s(raw_h, args.h_rows + q0 * kD, bytes, &sm.q_bar);
        tma_bulk_g2s(raw_t, args.t_rows + q0 * kD, bytes, &sm.q_bar);
        tma_bulk_g2s(raw_r, args.r_rows + q0 * kD, bytes, &sm.q_bar);
    }
    mbar_wait(&sm.q_bar, 0);
    for (int idx = tid; idx < kQT * kD; idx += kThreads) {
        const int q = idx >> 7, j = idx & (kD - 1);
        float *v0 = &sm.qv[q][0][0], *v1 = &sm.qv[q][1][0];
        if (q < nq) {
            const float *h = raw_h + q * kD, *t = raw_t + q * kD, *r = raw_r + q * kD;
            if (head_pred) fold_query<MODEL, true>(h, t, r, j, v0, v1);
            else fold_query<MODEL, false>(h, t, r, j, v0, v1);
        } else {
            v0[j] = 0.0f;
            v1[j] = 0.0f;
        }
    }
    if (tid < kQT) sm.st[tid] = (args.true_score && tid < nq) ? args.true_score[qout0 + tid] : 0.0f;
    __syncthreads();   // raw rows consumed: tile buffers may now be overwritten

    const long long ntiles = (args.n_local + kCT - 1) / kCT;

    if (warp == kConsumerWarps) {
        // ===================== producer warp =====================
        int it = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            const uint32_t use = (uint32_t)(it >> 1);
            mbar_wait(&sm.empty_bar[buf], (use & 1u) ^ 1u);
            const long long base = tile * kCT;
            const int rows = (int)min((long long)kCT, args.n_local - base);
            float *dst = &sm.ctile[buf][0][0];
            if (MODEL == BLP_MODEL_TRANSE && args.use_tma) {
                if (lane == 0) mbar_arrive_expect_tx(&sm.full_bar[buf], (uint32_t)rows * kD * 4u);
                __syncwarp();
                for (int row = lane; row < rows; row += 32)
                    tma_bulk_g2s(dst + row * kPitch, args.ent + (base + row) * kD, kD * 4u, &sm.full_bar[buf]);
            } else {
                // 2048 16-byte chunks per tile; a warp-wide load covers one 512-byte row
#pragma unroll 1
                for (int batch = 0; batch < kCT / 16; ++batch) {
                    float4 v[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        const int row = batch * 16 + u;
                        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (row < rows) v[u] = __ldg(reinterpret_cast<const float4 *>(args.ent + (base + row) * kD) + lane);
                    }
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        float *drow = dst + (batch * 16 + u) * kPitch;
                        if (MODEL == BLP_MODEL_TRANSE) {
                            *reinterpret_cast<float4 *>(drow + 4 * lane) = v[u];
                        } else {
                            drow[perm_pos<MODEL>(4 * lane + 0)] = v[u].x;
                            drow[perm_pos<MODEL>(4 * lane + 1)] = v[u].y;
                            drow[perm_pos<MODEL>(4 * lane + 2)] = v[u].z;
                            drow[perm_pos<MODEL>(4 * lane + 3)] = v[u].w;
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.full_bar[buf]);
            }
        }
    } else {
        // ===================== consumer warps =====================
        const float *qv = &sm.qv[warp * kTQ][0][0];
        float st[kTQ];
#pragma unroll
        for (int q = 0; q < kTQ; ++q) st[q] = sm.st[warp * kTQ + q];
        int cgt[kTQ], cge[kTQ];
#pragma unroll
        for (int q = 0; q < kTQ; ++q) cgt[q] = cge[q] = 0;

        int it = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            const uint32_t use = (uint32_t)(it >> 1);
            mbar_wait(&sm.full_bar[buf], use & 1u);
            float s[kTQ][kTC];
            if (head_pred) score_tile<MODEL, true>(&sm.ctile[buf][0][0], qv, lane, s);
            else score_tile<MODEL, false>(&sm.ctile[buf][0][0], qv, lane, s);
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty_bar[buf]);

            const long long base = tile * kCT;
            if (args.scores_out) {
#pragma unroll
                for (int q = 0; q < kTQ; ++q) {
                    const int ql = warp * kTQ + q;
                    if (ql < nq) {
#pragma unroll
                        for (int i = 0; i < kTC; ++i) {
                            const long long cand = base + lane + 32 * i;
                            if (cand < args.n_local) args.scores_out[(qout0 + ql) * args.ld_scores + cand] = s[q][i];
                        }
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < kTC; ++i) {
                    const bool valid = base + lane + 32 * i < args.n_local;
#pragma unroll
                    for (int q = 0; q < kTQ; ++q) {
                        cgt[q] += (valid && s[q][i] > st[q]) ? 1 : 0;
                        cge[q] += (valid && s[q][i] >= st[q]) ? 1 : 0;
                    }
                }
            }
        }
        if (!args.scores_out) {
#pragma unroll
            for (int q = 0; q < kTQ; ++q) {
                const int a = __reduce_add_sync(0xffffffffu, cgt[q]);
                const int c = __reduce_add_sync(0xffffffffu, cge[q]);
                const int ql = warp * kTQ + q;
                if (lane == 0 && ql < nq && it > 0) {
                    atomicAdd(&args.gt[qout0 + ql], a);
                    atomicAdd(&args.ge[qout0 + ql], c);
                }
            }
        }
    }
}

// ---- true-triple scores + counter reset ------------------------------------
// s_true for head query i and tail query i is the same number: the score of
// (h_i, r_i, t_i) (pred.gather(true_idx), utils.py:103).  One thread per triple.
__global__ void true_score_kernel(int model, const float *__restrict__ h_rows, const float *__restrict__ t_rows,
                                  const float *__restrict__ r_rows, long long b, int d, float *__restrict__ true_score,
                                  int *__restrict__ gt, int *__restrict__ ge) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b) return;
    const float s = score_exact_dyn(model, h_rows + i * d, t_rows + i * d, r_rows + i * d, d);
    true_score[i] = s;
    true_score[b + i] = s;
    gt[i] = 0; gt[b + i] = 0;
    ge[i] = 0; ge[b + i] = 0;
}

// ---- filtered ranks: sparse correction (train.py:159-167) -------------------
// One warp per query: re-score the query's filtered candidates that live in this
// shard and remove their contribution from the raw counts.
__global__ void filter_correct_kernel(int model, const float *__restrict__ ent, long long n_local, long long ent_offset,
                                      int d, const float *__restrict__ h_rows, const float *__restrict__ t_rows,
                                      const float *__restrict__ r_rows, long long b,
                                      const long long *__restrict__ indptr, const long long *__restrict__ idx,
                                      const float *__restrict__ true_score, const int *__restrict__ gt,
                                      const int *__restrict__ ge, int *__restrict__ gt_f, int *__restrict__ ge_f) {
    const long long q = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= 2 * b) return;
    const bool head_pred = q < b;
    const long long i = head_pred ? q : q - b;
    const float st = true_score[q];
    int cg = 0, ce = 0;
    if (indptr) {
        for (long long p = indptr[q] + lane; p < indptr[q + 1]; p += 32) {
            const long long row = idx[p] - ent_offset;
            if (row < 0 || row >= n_local) continue;
            const float *e = ent + row * d;
            const float s = head_pred ? score_exact_dyn(model, e, t_rows + i * d, r_rows + i * d, d)
                                      : score_exact_dyn(model, h_rows + i * d, e, r_rows + i * d, d);
            cg += s > st;
            ce += s >= st;
        }
    }
    cg = __reduce_add_sync(0xffffffffu, cg);
    ce = __reduce_add_sync(0xffffffffu, ce);
    if (lane == 0) {
        gt_f[q] = gt[q] - cg;
        ge_f[q] = ge[q] - ce;
    }
}

// ---- generic-width sweep (d != 128): one thread per (query, candidate) -----
__global__ void sweep_generic_kernel(int model, const float *__restrict__ ent, long long n_local, int d,
                                     const float *__restrict__ h_rows, const float *__restrict__ t_rows,
                                     const float *__restrict__ r_rows, long long b,
                                     const float *__restrict__ true_score, int *__restrict__ gt, int *__restrict__ ge) {
    const long long q = blockIdx.y;
    const bool head_pred = q < b;
    const long long i = head_pred ? q : q - b;
    const float st = true_score[q];
    int cg = 0, ce = 0;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n_local; c += (long long)gridDim.x * blockDim.x) {
        const float *e = ent + c * d;
        const float s = head_pred ? score_exact_dyn(model, e, t_rows + i * d, r_rows + i * d, d)
                                  : score_exact_dyn(model, h_rows + i * d, e, r_rows + i * d, d);
        cg += s > st;
        ce += s >= st;
    }
    cg = __reduce_add_sync(0xffffffffu, cg);
    ce = __reduce_add_sync(0xffffffffu, ce);
    if ((threadIdx.x & 31) == 0 && (cg | ce)) {
        atomicAdd(&gt[q], cg);
        atomicAdd(&ge[q], ce);
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
                                                              long long q, KValues kv, double *__restrict__ sums) {
    __shared__ double scratch[32][9];
    double acc[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) acc[j] = 0.0;
    for (long long i = threadIdx.x; i < q; i += blockDim.x) {
        const float avg = fmul((float)((long long)gt[i] + 1 + (long long)ge[i]), 0.5f);
        acc[0] += (double)__frcp_rn(avg);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < kv.nk && avg <= (float)kv.k[j]) acc[1 + j] += 1.0;
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
static int num_sms() {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
}

// number of candidate splits: fill the SMs with whole waves, keep tiles per CTA balanced
static int pick_splits(long long ntiles, long long groups, int sms) {
    long long best = 1;
    double best_eff = -1.0;
    const long long smax = ntiles < 64 ? ntiles : 64;
    for (long long s = 1; s <= smax; ++s) {
        const long long ctas = groups * s;
        const long long waves = (ctas + sms - 1) / sms;
        const long long tiles_per = (ntiles + s - 1) / s;
        // time ~ waves * tiles_per (+ a fixed per-CTA query staging cost of about a third of a tile)
        const double t = (double)waves * ((double)tiles_per + 0.35);
        const double eff = ((double)ntiles * groups / sms) / t;
        if (eff > best_eff * 1.001) { best_eff = eff; best = s; }
    }
    return (int)best;
}

template <int MODEL>
static int launch_sweep(const SweepArgs &a, cudaStream_t st) {
    const size_t smem = sizeof(SweepSmem);
    BLP_CUDA(cudaFuncSetAttribute(sweep_kernel<MODEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long ntiles = (a.n_local + kCT - 1) / kCT;
    const long long groups = (long long)a.groups_per_role * (a.roles == 3 ? 2 : 1);
    if (ntiles == 0 || groups == 0) return BLP_OK;
    if (groups > 65535) { set_error("too many query groups (%lld); split the sweep", groups); return BLP_EINVAL; }
    const int splits = pick_splits(ntiles, groups, num_sms());
    dim3 grid((unsigned)splits, (unsigned)groups);
    prof_begin(1, st);
    sweep_kernel<MODEL><<<grid, kThreads, smem, st>>>(a);
    prof_end(1, st);
    count_launch();
    BLP_CUDA(cudaGetLastError());
    return BLP_OK;
}

static int launch_sweep_dyn(int model, const SweepArgs &a, cudaStream_t st) {
    switch (model) {
    case BLP_MODEL_TRANSE: return launch_sweep<BLP_MODEL_TRANSE>(a, st);
    case BLP_MODEL_DISTMULT: return launch_sweep<BLP_MODEL_DISTMULT>(a, st);
    case BLP_MODEL_COMPLEX: return launch_sweep<BLP_MODEL_COMPLEX>(a, st);
    default: return launch_sweep<BLP_MODEL_SIMPLE>(a, st);
    }
}

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

static int env_use_tma() {
    const char *e = getenv("BLP_EVAL_PRODUCER");
    if (e && (e[0] == 'l' || e[0] == 'L')) return 0;   // "ldg"
    return 1;
}

}  // namespace blp

using namespace blp;

extern "C" int blp_eval_rank(int model, const float *ent, int64_t n_local, int64_t ent_offset, int d,
                             const float *h_rows, const float *t_rows, const float *r_rows, int64_t b,
                             const int64_t *filt_indptr, const int64_t *filt_idx, int32_t *gt, int32_t *ge,
                             int32_t *gt_f, int32_t *ge_f, float *true_score, void *stream) {
    reset_launch_count();
    int rc = check_model_dim(model, d);
    if (rc) return rc;
    if (b < 0 || n_local < 0) { set_error("negative size"); return BLP_EINVAL; }
    if (b == 0) return BLP_OK;
    if (!h_rows || !t_rows || !r_rows || !gt || !ge || !true_score || (n_local > 0 && !ent)) {
        set_error("null pointer argument");
        return BLP_EINVAL;
    }
    if (filt_indptr && (!gt_f || !ge_f || !filt_idx)) { set_error("filters given without gt_f/ge_f/filt_idx"); return BLP_EINVAL; }
    if (n_local >= (1ll << 31)) { set_error("n_local too large for int32 counters"); return BLP_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;

    true_score_kernel<<<(unsigned)((b + 127) / 128), 128, 0, st>>>(model, h_rows, t_rows, r_rows, b, d, true_score, gt, ge);
    count_launch();
    BLP_CUDA(cudaGetLastError());

    if (n_local > 0) {
        if (d == kD && aligned16(ent) && aligned16(h_rows) && aligned16(t_rows) && aligned16(r_rows)) {
            SweepArgs a{};
            a.ent = ent; a.n_local = n_local; a.h_rows = h_rows; a.t_rows = t_rows; a.r_rows = r_rows; a.b = b;
            a.true_score = true_score; a.gt = gt; a.ge = ge; a.scores_out = nullptr; a.ld_scores = 0;
            a.roles = 3; a.groups_per_role = (int)((b + kQT - 1) / kQT); a.use_tma = env_use_tma();
            // the grid's y extent is limited to 65535 groups: chunk very long sweeps
            const long long max_b = 32000ll * kQT;
            if (b > max_b) { set_error("b > %lld: split the sweep into several calls", max_b); return BLP_EINVAL; }
            rc = launch_sweep_dyn(model, a, st);
            if (rc) return rc;
        } else {
            const int threads = 256;
            long long bx = (n_local + threads - 1) / threads;
            if (bx > 1024) bx = 1024;
            if (2 * b > 65535) { set_error("generic-width sweep: b too large, split the sweep"); return BLP_EINVAL; }
            dim3 grid((unsigned)bx, (unsigned)(2 * b));
            sweep_generic_kernel<<<grid, threads, 0, st>>>(model, ent, n_local, d, h_rows, t_rows, r_rows, b, true_score, gt, ge);
            count_launch();
            BLP_CUDA(cudaGetLastError());
        }
    }
    if (gt_f && ge_f) {
        const long long warps = 2 * b;
        const int threads = 128;
        const long long blocks = (warps * 32 + threads - 1) / threads;
        filter_correct_kernel<<<(unsigned)blocks, threads, 0, st>>>(model, ent, n_local, ent_offset, d, h_rows, t_rows,
                                                                    r_rows, b, (const long long *)filt_indptr,
                                                                    (const long long *)filt_idx, true_score, gt, ge, gt_f, ge_f);
        count_launch();
        BLP_CUDA(cudaGetLastError());
    }
    return BLP_OK;
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
    if (d == kD && (cand_h || cand_t) && C >= 32 && A <= 32000ll * kQT && aligned16(heads) && aligned16(tails) && aligned16(rels)) {
        SweepArgs a{};
        a.ent = cand_h ? heads : tails; a.n_local = C;
        // the unused query operand aliases a valid row block so the TMA staging reads defined memory
        a.h_rows = cand_h ? tails : heads; a.t_rows = cand_h ? tails : heads; a.r_rows = rels;
        if (cand_h) a.t_rows = tails; else a.h_rows = heads;
        a.b = A; a.true_score = nullptr; a.gt = nullptr; a.ge = nullptr; a.scores_out = out; a.ld_scores = C;
        a.roles = cand_h ? 1 : 2; a.groups_per_role = (int)((A + kQT - 1) / kQT); a.use_tma = env_use_tma();
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
    reset_launch_count();
    if (q < 0 || nk < 0 || nk > 8) { set_error("bad q or nk (nk <= 8)"); return BLP_EINVAL; }
    if (!sums || (q > 0 && (!gt || !ge)) || (nk > 0 && !k_values_host)) { set_error("null pointer argument"); return BLP_EINVAL; }
    KValues kv{};
    kv.nk = nk;
    for (int i = 0; i < nk; ++i) kv.k[i] = k_values_host[i];
    metrics_reduce_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(gt, ge, q, kv, sums);
    count_launch();
    BLP_CUDA(cudaGetLastError());
    return BLP_OK;
}
