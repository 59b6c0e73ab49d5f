// blp_sweep.cu -- the fused full-entity scoring + rank-counting kernel (a1-a4, a10, a11).
//
// Replaces train.py:141-157 + utils.py:103-105 for d == 128 (the width of every BLP script): for every
// query (a test triple with its head or its tail removed) score ALL candidate entity rows and reduce the
// scores to two integers, gt = #{s_j > s_true} and ge = #{s_j >= s_true}.  Neither the (B, N, D)
// broadcast nor the (2B, N) score matrix is materialised.
//
// Layout
//   - PERSISTENT grid, one CTA per SM, 8 consumer warps + 1 producer warp.  The work is the list of
//     (triple group, candidate tile) pairs split evenly over the CTAs; a CTA crosses at most a few group
//     boundaries and re-folds its queries at each.
//   - a triple group carries BOTH roles: every consumer thread predicts the heads of its triples (the
//     candidate row plays `heads`, train.py:146) with half of its query pairs and their tails
//     (train.py:147) with the other half, so all warps do identical work, consume tiles in step, and each
//     candidate tile is fetched once for both roles.
//   - a slot is the 2 * TQP queries of one warp; a consumer thread owns a (TQP query pairs) x (TC
//     candidates) register tile.  Large batches use TQP = 4, TC = 4 (32 triples = 64 queries per group,
//     every warp covers all 128 tile rows).  Small batches (the reference's Wikidata5M eval_batch_size = 2) shrink
//     TQP / TC and split the tile rows over 4 / TC warps per slot instead, so no lane-ops are spent on
//     padding queries and the kernel becomes HBM bound.
//   - query rows (h, t, r) are folded into <= 2 operand vectors per query (e.g. u = fl(h + r) for TransE
//     tail prediction), stored in shared memory in *processing order*, interleaved in PAIRS of queries so
//     one 64-bit register pair feeds a packed FADD2 / FFMA2 (two queries per issue slot);
//   - candidate tiles (128 rows) are multi-buffered in shared memory with a 132-float pitch, so a lane
//     that walks one row with 128-bit loads never conflicts; the producer warp fills them either with
//     TMA bulk copies (natural order: TransE) or with coalesced 128-bit global loads + a permuting store
//     (ATen sum order: DistMult / ComplEx / SimplE); mbarrier full / empty pairs per buffer;
//   - every (query, candidate) replays the reference's exact fp32 operation order
//     and is compared against the true-triple score.
// The kernel is FP32-pipe bound for full groups (2-3 lane-ops per (query, candidate, dim) against 4/Q
// bytes) and HBM bound for the small-batch shapes; see DESIGN.md.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "blp_sweep.h"

// unroll factors of the TransE position loop (tuning: instruction-cache footprint vs. scheduling freedom)
#ifndef BLP_UNROLL_CC
#define BLP_UNROLL_CC 8
#endif
#ifndef BLP_UNROLL_CC_SHARED
#define BLP_UNROLL_CC_SHARED 8
#endif

namespace blp {

constexpr int kPitch = 132;                   // smem row pitch (floats) of the padded (bilinear) tile layout
constexpr int kCW = 8;                        // consumer warps (default; Cfg<4, 2, 12> / Cfg<4, 2, 16> were measured and are slower)
constexpr int kCT = 128;                      // candidate rows per tile
constexpr int kSmemBudget = 227 * 1024 - 1024;

// TS_ (tile split): every slot ranks the SAME triples and the slots take turns on the candidate tiles (slot = tile index
// mod NS), instead of every slot ranking its own triples on every tile.  Cfg<2, 1, 8, true> is the pass-size-2 shape
// (Wikidata5M eval batch): a thread holds the head AND the tail pair of the two triples for one row -- two independent
// accumulation chains and one candidate load for both -- where Cfg<1, 1> gives a thread a single chain.
template <int TQP_, int TC_, int CW_ = kCW, bool TS_ = false>
struct Cfg {
    static constexpr bool TS = TS_;
    static constexpr int TQP = TQP_;          // query pairs per consumer thread
    static constexpr int TC = TC_;            // candidates per consumer thread
    static constexpr int CW = CW_;            // consumer warps of the CTA
    static constexpr int RS = 4 / TC_;        // warps that share one slot (they split the 128 tile rows)
    static constexpr int NS = CW_ / RS;       // slots per CTA
    static constexpr int SQ = 2 * TQP_;       // queries per slot
    static constexpr int NQ = NS * SQ;        // queries per CTA
    static constexpr int QV_BYTES = NS * TQP_ * 2 * kD * 2 * 4;
};

// Staged tile layouts (one lane walks one candidate row with 128-bit loads; both are conflict-free):
//   TransE    natural element order, written by TMA tensor copies as four 32-float column blocks of
//             [128 rows][32 floats] with the 128-byte swizzle (16-byte chunk c of row r sits at chunk c ^ (r & 7));
//   bilinear  ATen sum order (perm_pos), written by the producer warp, rows padded to 132 floats.
template <int MODEL>
struct TileLayout {
    static constexpr bool kSwizzled = MODEL == BLP_MODEL_TRANSE;
    static constexpr int kFloats = kSwizzled ? kCT * kD : kCT * kPitch;
    // 16-byte chunk holding staged positions [4 * pos4, 4 * pos4 + 4) of tile row `row`
    __device__ static __forceinline__ int chunk(int row, int pos4) {
        if (kSwizzled) return (pos4 >> 3) * (kCT * 32) + row * 32 + (((pos4 & 7) ^ (row & 7)) << 2);
        return row * kPitch + 4 * pos4;
    }
};

template <int MODEL, class C>
struct Stages {
    // per query: 128 term floats (in-kernel true scores) + three row pointers + flags
    static constexpr int kPerWarp = (C::NQ + C::CW - 1) / C::CW;            // queries per warp (ql = warp + CW * u)
    static constexpr int kRoundMax = C::CW > 8 ? 2 : 4;                     // (16 warps: 2, the term rows must fit next to the tiles)
    static constexpr int kRound = kPerWarp < kRoundMax ? kPerWarp : kRoundMax;   // queries per warp and round
    static constexpr int kTermRows = C::CW * kRound;
    static constexpr int kPrologueBytes = kTermRows * kD * 4 + C::NQ * (3 * 8 + 3 * 4 + 4);
    static constexpr int ST = (C::QV_BYTES + kPrologueBytes + 3 * TileLayout<MODEL>::kFloats * 4 <= kSmemBudget) ? 3 : 2;   // tile buffers
};

template <int MODEL, class C>
struct __align__(1024) SweepSmem {
    static constexpr int ST = Stages<MODEL, C>::ST;
    float ctile[ST][TileLayout<MODEL>::kFloats];   // candidate tiles (1024-byte aligned for the TMA swizzle)
    float qv[C::NS * C::TQP][2][kD][2];       // [query pair][operand vector][position][half], processing order
    float tm[Stages<MODEL, C>::kTermRows][kD];   // per-position terms of the true-triple scores: 4 queries per warp and round
    const float *rowp[C::NQ][3];              // h / t / r row of every query of the current group
    int rowok[C::NQ][3];                      // index inside its table (h / t / r)
    float st[C::NQ];                          // true-triple scores of the current group's queries
    // tile split: a parity wait is only sound for a waiter that observes EVERY phase of its barrier, and a slot sees only
    // every NS-th tile -- so each (slot, buffer) combination gets its own `full` barrier (tile it -> it % (ST * NS)); the
    // producer is the only waiter of the `empty` barriers and sees all of their phases
    static constexpr int kFull = ST * (C::TS ? C::NS : 1);
    uint64_t full_bar[kFull];
    uint64_t empty_bar[ST];
};

// per-lane view of a staged tile: rows row + 32 * i, i < TC.  All of a lane's rows share row & 7, so the
// swizzle is one XOR of the (compile-time) chunk index with a per-lane constant.
template <int MODEL>
struct TileView {
    const float *lane_base;   // tile base + this lane's first row
    int xr;                   // (row & 7) << 2 for the swizzled layout
    __device__ __forceinline__ TileView(const float *tile, int row)
        : lane_base(tile + (TileLayout<MODEL>::kSwizzled ? row * 32 : row * kPitch)), xr((row & 7) << 2) {}
    // 16-byte chunk holding staged positions [4 * pos4, 4 * pos4 + 4) of row (row + 32 * i)
    __device__ __forceinline__ float4 load(int i, int pos4) const {
        if (TileLayout<MODEL>::kSwizzled)
            return *reinterpret_cast<const float4 *>(lane_base + i * (32 * 32) + (pos4 >> 3) * (kCT * 32) + (((pos4 & 7) << 2) ^ xr));
        return *reinterpret_cast<const float4 *>(lane_base + i * (32 * kPitch) + 4 * pos4);
    }
};

// position of natural element j inside a staged row
template <int MODEL>
__host__ __device__ __forceinline__ constexpr int perm_pos(int j) {
    if (MODEL == BLP_MODEL_TRANSE) return j;
    if (MODEL == BLP_MODEL_DISTMULT) return (j & 7) * 16 + ((j >> 3) & 3) * 4 + (j >> 5);
    return (j >> 6) * 64 + (j & 7) * 8 + ((j >> 3) & 3) * 2 + ((j & 63) >> 5);
}

// ---- query-side folding -----------------------------------------------------
// HEAD_PRED: the candidate row plays `heads` (train.py:146); otherwise `tails` (train.py:147).
// qp points at qv[pair][0][0][half]; operand vector v, position pos live at qp[(v * kD + pos) * 2].
template <int MODEL, bool HEAD_PRED>
__device__ __forceinline__ void fold_query(const float *__restrict__ h, const float *__restrict__ t,
                                           const float *__restrict__ r, int j, float *__restrict__ qp) {
    const int p = perm_pos<MODEL>(j);
    auto put = [&](int v, int pos, float x) { qp[(v * kD + pos) * 2] = x; };
    if (MODEL == BLP_MODEL_TRANSE || MODEL == BLP_MODEL_DISTMULT) {
        if (HEAD_PRED) {
            put(0, p, r[j]);
            put(1, p, t[j]);
        } else {
            put(0, p, (MODEL == BLP_MODEL_TRANSE) ? fadd(h[j], r[j]) : fmul(h[j], r[j]));
            put(1, p, 0.0f);
        }
    } else if (MODEL == BLP_MODEL_COMPLEX) {
        if (HEAD_PRED) {
            put(0, p, r[j]);
            put(1, p, t[j]);
        } else if (j < 64) {
            const float hr = h[j], hi = h[64 + j], rr = r[j], ri = r[64 + j];
            put(0, p, fmul(rr, hr));        // A
            put(0, 64 + p, fmul(rr, hi));   // B
            put(1, p, fmul(ri, hr));        // C
            put(1, 64 + p, fmul(ri, hi));   // D
        }
    } else {  // SIMPLE
        if (j < 64) {
            if (HEAD_PRED) {             // candidate = (hh, ht)
                put(0, p, r[j]);                          // ra
                put(0, 64 + p, t[64 + j]);                // tt
                put(1, p, fmul(t[j], r[64 + j]));         // th * rb
                put(1, 64 + p, 0.0f);
            } else {                     // candidate = (th, tt)
                put(0, p, fmul(h[j], r[j]));              // hh * ra
                put(0, 64 + p, r[64 + j]);                // rb
                put(1, p, h[64 + j]);                     // ht
                put(1, 64 + p, 0.0f);
            }
        }
    }
}

// The same folding from values already in registers: h0 = h[j], t0 = t[j], r0 = r[j]; for the halves models j < 64 and
// h1 = h[64 + j], t1 = t[64 + j], r1 = r[64 + j].
template <int MODEL, bool HEAD_PRED>
__device__ __forceinline__ void fold_vals(float h0, float t0, float r0, float h1, float t1, float r1, int j,
                                          float *__restrict__ qp) {
    const int p = perm_pos<MODEL>(j);
    auto put = [&](int v, int pos, float x) { qp[(v * kD + pos) * 2] = x; };
    if (MODEL == BLP_MODEL_TRANSE || MODEL == BLP_MODEL_DISTMULT) {
        if (HEAD_PRED) {
            put(0, p, r0);
            put(1, p, t0);
        } else {
            put(0, p, (MODEL == BLP_MODEL_TRANSE) ? fadd(h0, r0) : fmul(h0, r0));
            put(1, p, 0.0f);
        }
    } else if (MODEL == BLP_MODEL_COMPLEX) {
        if (HEAD_PRED) {                 // v0 = (rr | ri), v1 = (tr | ti)
            put(0, p, r0); put(0, 64 + p, r1);
            put(1, p, t0); put(1, 64 + p, t1);
        } else {
            put(0, p, fmul(r0, h0));        // A
            put(0, 64 + p, fmul(r0, h1));   // B
            put(1, p, fmul(r1, h0));        // C
            put(1, 64 + p, fmul(r1, h1));   // D
        }
    } else {  // SIMPLE
        if (HEAD_PRED) {                 // candidate = (hh, ht)
            put(0, p, r0);                            // ra
            put(0, 64 + p, t1);                       // tt
            put(1, p, fmul(t0, r1));                  // th * rb
            put(1, 64 + p, 0.0f);
        } else {                         // candidate = (th, tt)
            put(0, p, fmul(h0, r0));                  // hh * ra
            put(0, 64 + p, r1);                       // rb
            put(1, p, h1);                            // ht
            put(1, 64 + p, 0.0f);
        }
    }
}

// one per-position term of the true-triple score, operation order of models.py:222-248
template <int MODEL>
__device__ __forceinline__ float true_term(float h0, float t0, float r0, float h1, float t1, float r1) {
    if (MODEL == BLP_MODEL_TRANSE) return fabsf(fsub(fadd(h0, r0), t0));
    if (MODEL == BLP_MODEL_DISTMULT) return fmul(fmul(h0, r0), t0);
    if (MODEL == BLP_MODEL_COMPLEX) {        // (hr, hi) = (h0, h1), (tr, ti) = (t0, t1), (rr, ri) = (r0, r1)
        float p = fadd(fmul(fmul(r0, h0), t0), fmul(fmul(r0, h1), t1));
        p = fadd(p, fmul(fmul(r1, h0), t1));
        return fsub(p, fmul(fmul(r1, h1), t0));
    }
    return fadd(fmul(fmul(h0, r0), t1), fmul(fmul(t0, r1), h1));   // SimplE: hh ra tt + th rb ht
}

// ---- per-position arithmetic on query PAIRS (f2 = two queries, same candidate) ----
__device__ __forceinline__ f2 transe_step(bool HEAD_PRED, f2 acc, float e, f2 v0, f2 v1) {
    const f2 x = HEAD_PRED ? sub2(add2(dup2(e), v0), v1) : sub2(v0, dup2(e));
    return add2(acc, abs2(x));
}
__device__ __forceinline__ f2 distmult_term(bool HEAD_PRED, float e, f2 v0, f2 v1, f2 nz) {
    return HEAD_PRED ? mul2(mul2(dup2(e), v0, nz), v1, nz) : mul2(v0, dup2(e), nz);
}
template <int MODEL>
__device__ __forceinline__ f2 halves_term(bool HEAD_PRED, float elo_, float ehi_, f2 v0lo, f2 v0hi, f2 v1lo, f2 v1hi, f2 nz) {
    const f2 elo = dup2(elo_), ehi = dup2(ehi_);
    if (MODEL == BLP_MODEL_COMPLEX) {
        if (HEAD_PRED) {  // e = (hr, hi); v0 = (rr, ri); v1 = (tr, ti)
            f2 p = add2(mul2(mul2(v0lo, elo, nz), v1lo, nz), mul2(mul2(v0lo, ehi, nz), v1hi, nz));
            p = add2(p, mul2(mul2(v0hi, elo, nz), v1hi, nz));
            return sub2(p, mul2(mul2(v0hi, ehi, nz), v1lo, nz));
        } else {          // e = (tr, ti); v0 = (A, B); v1 = (C, D)
            f2 p = add2(mul2(v0lo, elo, nz), mul2(v0hi, ehi, nz));
            p = add2(p, mul2(v1lo, ehi, nz));
            return sub2(p, mul2(v1hi, elo, nz));
        }
    } else {
        if (HEAD_PRED) return add2(mul2(mul2(elo, v0lo, nz), v0hi, nz), mul2(v1lo, ehi, nz));   // (hh*ra)*tt + (th*rb)*ht
        return add2(mul2(v0lo, ehi, nz), mul2(mul2(elo, v0hi, nz), v1lo, nz));                  // (hh*ra)*tt + (th*rb)*ht
    }
}

__device__ __forceinline__ float4 lds128(const float *p) { return *reinterpret_cast<const float4 *>(p); }

// four consecutive staged positions of one operand vector of one query pair
struct Q4 {
    f2 x, y, z, w;
};
__device__ __forceinline__ Q4 ldq4(const float *qv, int q, int v, int off) {
    const float *p = qv + ((q * 2 + v) * kD + off) * 2;
    const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(p);
    const ulonglong2 b = *reinterpret_cast<const ulonglong2 *>(p + 4);
    Q4 r;
    r.x = a.x; r.y = a.y; r.z = b.x; r.w = b.y;
    return r;
}
__device__ __forceinline__ Q4 q4_zero() {
    Q4 r;
    r.x = r.y = r.z = r.w = 0ull;
    return r;
}

// role of query pair q inside a thread's register tile
constexpr int kRoleHead = 0, kRoleTail = 1, kRoleMixed = 2;   // mixed: first half of the pairs head, second half tail
template <int ROLE, int TQP>
__host__ __device__ __forceinline__ constexpr bool pair_is_head(int q) {
    return ROLE == kRoleHead || (ROLE == kRoleMixed && q < TQP / 2);
}

// Scores of a (TQP query pairs) x (TC candidates) register tile against the staged rows tv.row + 32 * i.
// qv = this warp's first query pair.  On return s[q][i] holds the final bilinear scores of queries 2q, 2q+1,
// or the L1 distance (the caller flips the sign) for TransE.  Loop nests are ordered (pair, position,
// candidate) so consecutive packed instructions belong to different accumulation chains.
//
// SHARED_R (TransE, mixed role): the thread's TQP head-prediction queries belong to triples with the SAME
// relation (test triples sorted by relation), so w = fl(e + r) -- the first rounding of models.py:223 when the
// candidate plays `heads` -- is computed once per (candidate, position) as a scalar and reused by every head
// pair: 1 lane-op instead of 2 per pair.  Per query the operations and their order are unchanged (same bits).
template <int MODEL, int ROLE, int TQP, int TC, bool SHARED_R = false>
__device__ __forceinline__ void score_tile(const TileView<MODEL> tv, const float *__restrict__ qv, f2 nz,
                                           f2 (&s)[TQP][TC]) {
#define HEAD_PRED (pair_is_head<ROLE, TQP>(q))
#pragma unroll
    for (int q = 0; q < TQP; ++q)
#pragma unroll
        for (int i = 0; i < TC; ++i) s[q][i] = 0ull;

    if (MODEL == BLP_MODEL_TRANSE) {
        // strictly sequential L1 accumulation, natural order
        // the kernel holds two copies of this loop nest (shared / per-query relation); only one of them is hot at any
        // time (the choice is per triple group), so both can be unrolled fully without instruction-cache thrash
        constexpr int kUnrollCC = SHARED_R ? BLP_UNROLL_CC_SHARED : BLP_UNROLL_CC;
#pragma unroll 1
        for (int cb = 0; cb < 4; ++cb)
#pragma unroll kUnrollCC
        for (int cc = 0; cc < 8; ++cc) {
            const int c4 = cb * 8 + cc;
            float e[TC][4];
#pragma unroll
            for (int i = 0; i < TC; ++i) {
                const float4 v = tv.load(i, c4);
                e[i][0] = v.x; e[i][1] = v.y; e[i][2] = v.z; e[i][3] = v.w;
            }
            constexpr int kHeadPairs = (SHARED_R && ROLE == kRoleMixed) ? TQP / 2 : 0;
            if (kHeadPairs > 0) {
                // shared relation: r is the same in both halves of every head pair; take it from pair 0
                const Q4 r4 = ldq4(qv, 0, 0, 4 * c4);
                const f2 rv[4] = {r4.x, r4.y, r4.z, r4.w};
                f2 tv4[kHeadPairs > 0 ? kHeadPairs : 1][4];
#pragma unroll
                for (int q = 0; q < kHeadPairs; ++q) {
                    const Q4 b = ldq4(qv, q, 1, 4 * c4);
                    tv4[q][0] = b.x; tv4[q][1] = b.y; tv4[q][2] = b.z; tv4[q][3] = b.w;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float r_lo, r_hi;
                    unpack2(rv[k], r_lo, r_hi);
#pragma unroll
                    for (int i = 0; i < TC; ++i) {
                        const f2 w = dup2(fadd(e[i][k], r_lo));
#pragma unroll
                        for (int q = 0; q < kHeadPairs; ++q) s[q][i] = add2(s[q][i], abs2(sub2(w, tv4[q][k])));
                    }
                }
            }
#pragma unroll
            for (int q = kHeadPairs; q < TQP; ++q) {
                const Q4 a = ldq4(qv, q, 0, 4 * c4);
                const Q4 b = HEAD_PRED ? ldq4(qv, q, 1, 4 * c4) : q4_zero();
                const f2 av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int i = 0; i < TC; ++i) s[q][i] = transe_step(HEAD_PRED, s[q][i], e[i][k], av[k], bv[k]);
            }
        }
    } else if (MODEL == BLP_MODEL_DISTMULT) {
        // staged order: pos = l*16 + a*4 + k  <->  j = 32k + 8a + l; one 16-byte chunk = one (l, a) chain
        for (int l = 0; l < 8; ++l) {
            f2 c[TQP][TC];
#pragma unroll
            for (int q = 0; q < TQP; ++q)
#pragma unroll
                for (int i = 0; i < TC; ++i) c[q][i] = 0ull;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int off = l * 16 + a * 4;
                float e[TC][4];
#pragma unroll
                for (int i = 0; i < TC; ++i) {
                    const float4 v = tv.load(i, off >> 2);
                    e[i][0] = v.x; e[i][1] = v.y; e[i][2] = v.z; e[i][3] = v.w;
                }
                constexpr int kHeadPairs = (SHARED_R && ROLE == kRoleMixed) ? TQP / 2 : 0;
                if (kHeadPairs > 0) {
                    // shared relation (relation-aligned order): fl(e * r) -- the first rounding of models.py:227 when the
                    // candidate plays `heads` -- once per (candidate, position) for all head pairs of the slot
                    const Q4 r4 = ldq4(qv, 0, 0, off);
                    const f2 rv[4] = {r4.x, r4.y, r4.z, r4.w};
                    f2 w[TC][4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float r_lo, r_hi;
                        unpack2(rv[k], r_lo, r_hi);
#pragma unroll
                        for (int i = 0; i < TC; ++i) w[i][k] = dup2(fmul(e[i][k], r_lo));
                    }
#pragma unroll
                    for (int q = 0; q < kHeadPairs; ++q) {
                        const Q4 v1 = ldq4(qv, q, 1, off);
                        const f2 v1v[4] = {v1.x, v1.y, v1.z, v1.w};
                        f2 chain[TC];
#pragma unroll
                        for (int i = 0; i < TC; ++i) chain[i] = mul2(w[i][0], v1v[0], nz);
#pragma unroll
                        for (int k = 1; k < 4; ++k)
#pragma unroll
                            for (int i = 0; i < TC; ++i) chain[i] = add2(chain[i], mul2(w[i][k], v1v[k], nz));
#pragma unroll
                        for (int i = 0; i < TC; ++i) c[q][i] = add2(c[q][i], chain[i]);
                    }
                }
#pragma unroll
                for (int q = kHeadPairs; q < TQP; ++q) {
                    const Q4 v0 = ldq4(qv, q, 0, off);
                    const Q4 v1 = HEAD_PRED ? ldq4(qv, q, 1, off) : q4_zero();
                    const f2 v0v[4] = {v0.x, v0.y, v0.z, v0.w}, v1v[4] = {v1.x, v1.y, v1.z, v1.w};
                    f2 chain[TC];
#pragma unroll
                    for (int i = 0; i < TC; ++i) chain[i] = distmult_term(HEAD_PRED, e[i][0], v0v[0], v1v[0], nz);
#pragma unroll
                    for (int k = 1; k < 4; ++k)
#pragma unroll
                        for (int i = 0; i < TC; ++i)
                            chain[i] = add2(chain[i], distmult_term(HEAD_PRED, e[i][k], v0v[k], v1v[k], nz));
#pragma unroll
                    for (int i = 0; i < TC; ++i) c[q][i] = add2(c[q][i], chain[i]);
                }
            }
#pragma unroll
            for (int q = 0; q < TQP; ++q)
#pragma unroll
                for (int i = 0; i < TC; ++i) s[q][i] = add2(s[q][i], c[q][i]);
        }
    } else {
        // halves; staged order per half: pos = l*8 + a*2 + k; one chunk = two (l, a) chains of length 2
        for (int l = 0; l < 8; ++l) {
            f2 c[TQP][TC];
#pragma unroll
            for (int q = 0; q < TQP; ++q)
#pragma unroll
                for (int i = 0; i < TC; ++i) c[q][i] = 0ull;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int off = l * 8 + hh * 4;
                float4 elo[TC], ehi[TC];
#pragma unroll
                for (int i = 0; i < TC; ++i) {
                    elo[i] = tv.load(i, off >> 2);
                    ehi[i] = tv.load(i, (off >> 2) + 16);
                }
                constexpr int kHeadPairsH = (SHARED_R && ROLE == kRoleMixed) ? TQP / 2 : 0;
                if (kHeadPairsH > 0) {
                    // shared relation: the products of the candidate with the relation -- ComplEx rr*hr, rr*hi, ri*hr, ri*hi
                    // (models.py:235-238), SimplE hh*ra (models.py:247) -- once per (candidate, position) for all head pairs
                    const Q4 ra4 = ldq4(qv, 0, 0, off);
                    const Q4 rb4 = (MODEL == BLP_MODEL_COMPLEX) ? ldq4(qv, 0, 0, 64 + off) : q4_zero();
                    const f2 rav[4] = {ra4.x, ra4.y, ra4.z, ra4.w}, rbv[4] = {rb4.x, rb4.y, rb4.z, rb4.w};
                    float rr[4], ri[4];
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        float dummy;
                        unpack2(rav[x], rr[x], dummy);
                        unpack2(rbv[x], ri[x], dummy);
                    }
                    Q4 hv1lo[kHeadPairsH > 0 ? kHeadPairsH : 1], hv1hi[kHeadPairsH > 0 ? kHeadPairsH : 1];
#pragma unroll
                    for (int q = 0; q < kHeadPairsH; ++q) {
                        // ComplEx: v1 = (tr | ti); SimplE: tt lives in v0's second half, th*rb in v1's first half
                        hv1lo[q] = ldq4(qv, q, 1, off);
                        hv1hi[q] = (MODEL == BLP_MODEL_COMPLEX) ? ldq4(qv, q, 1, 64 + off) : ldq4(qv, q, 0, 64 + off);
                    }
#pragma unroll
                    for (int i = 0; i < TC; ++i) {
                        const float el[4] = {elo[i].x, elo[i].y, elo[i].z, elo[i].w}, eh[4] = {ehi[i].x, ehi[i].y, ehi[i].z, ehi[i].w};
                        f2 wa[4], wb[4], wc[4], wd[4];
#pragma unroll
                        for (int x = 0; x < 4; ++x) {
                            wa[x] = dup2(fmul(rr[x], el[x]));                   // ComplEx rr*hr, SimplE hh*ra
                            if (MODEL == BLP_MODEL_COMPLEX) {
                                wb[x] = dup2(fmul(rr[x], eh[x]));               // rr*hi
                                wc[x] = dup2(fmul(ri[x], el[x]));               // ri*hr
                                wd[x] = dup2(fmul(ri[x], eh[x]));               // ri*hi
                            } else {
                                wb[x] = dup2(eh[x]);                            // ht
                                wc[x] = wd[x] = 0ull;
                            }
                        }
#pragma unroll
                        for (int q = 0; q < kHeadPairsH; ++q) {
                            const f2 t_lo[4] = {hv1lo[q].x, hv1lo[q].y, hv1lo[q].z, hv1lo[q].w};
                            const f2 t_hi[4] = {hv1hi[q].x, hv1hi[q].y, hv1hi[q].z, hv1hi[q].w};
                            f2 pp[4];
#pragma unroll
                            for (int x = 0; x < 4; ++x) {
                                if (MODEL == BLP_MODEL_COMPLEX) {
                                    f2 p = add2(mul2(wa[x], t_lo[x], nz), mul2(wb[x], t_hi[x], nz));
                                    p = add2(p, mul2(wc[x], t_hi[x], nz));
                                    pp[x] = sub2(p, mul2(wd[x], t_lo[x], nz));
                                } else {   // (hh*ra)*tt + (th*rb)*ht : t_hi = tt, t_lo = th*rb
                                    pp[x] = add2(mul2(wa[x], t_hi[x], nz), mul2(t_lo[x], wb[x], nz));
                                }
                            }
                            const f2 cc = add2(c[q][i], add2(pp[0], pp[1]));
                            c[q][i] = add2(cc, add2(pp[2], pp[3]));
                        }
                    }
                }
#pragma unroll
                for (int q = kHeadPairsH; q < TQP; ++q) {
                    const Q4 v0lo = ldq4(qv, q, 0, off);
                    const Q4 v0hi = ldq4(qv, q, 0, 64 + off);
                    const Q4 v1lo = ldq4(qv, q, 1, off);
                    const Q4 v1hi = (MODEL == BLP_MODEL_COMPLEX) ? ldq4(qv, q, 1, 64 + off) : q4_zero();
#pragma unroll
                    for (int i = 0; i < TC; ++i) {
                        const f2 p0 = halves_term<MODEL>(HEAD_PRED, elo[i].x, ehi[i].x, v0lo.x, v0hi.x, v1lo.x, v1hi.x, nz);
                        const f2 p1 = halves_term<MODEL>(HEAD_PRED, elo[i].y, ehi[i].y, v0lo.y, v0hi.y, v1lo.y, v1hi.y, nz);
                        const f2 p2 = halves_term<MODEL>(HEAD_PRED, elo[i].z, ehi[i].z, v0lo.z, v0hi.z, v1lo.z, v1hi.z, nz);
                        const f2 p3 = halves_term<MODEL>(HEAD_PRED, elo[i].w, ehi[i].w, v0lo.w, v0hi.w, v1lo.w, v1hi.w, nz);
                        const f2 cc = add2(c[q][i], add2(p0, p1));
                        c[q][i] = add2(cc, add2(p2, p3));
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < TQP; ++q)
#pragma unroll
                for (int i = 0; i < TC; ++i) s[q][i] = add2(s[q][i], c[q][i]);
        }
        if (MODEL == BLP_MODEL_SIMPLE) {
#pragma unroll
            for (int q = 0; q < TQP; ++q)
#pragma unroll
                for (int i = 0; i < TC; ++i) s[q][i] = mul2(s[q][i], dup2(0.5f), nz);
        }
    }
}
#undef HEAD_PRED

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define BLP_TS(slot)                                                                                  \
    do {                                                                                              \
        if (args.dbg && tid == 0) args.dbg[(size_t)blockIdx.x * 16 + (slot)] = globaltimer_ns();      \
    } while (0)

// Programmatic dependent launch (blp_plan_set_overlap): the next launch of the stream may start while this grid is
// still running (its CTAs take the SMs as ours exit) and blocks in griddep_wait() until this grid has completed and
// flushed.  Both are no-ops for launches without the attribute.
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <int CW>
__device__ __forceinline__ void consumer_bar_sync_n() {
    asm volatile("bar.sync 1, %0;" ::"n"(CW * 32) : "memory");
}
#define consumer_bar_sync() consumer_bar_sync_n<C::CW>()

// How the queries of a triple group map onto slots (a slot = the SQ queries of one warp-set):
//   ROLES == 3, TQP >= 2  "mixed": every slot takes TQP triples; its first TQP queries (pairs [0, TQP/2))
//                          predict their heads, the other TQP queries predict the tails of the SAME triples,
//                          so all warps do identical work and consume tiles in step;
//   ROLES == 3, TQP == 1  "split": a thread holds one pair, so the first half of the slots predicts heads
//                          and the second half tails (2 triples per group);
//   ROLES == 1 / 2         single role (the score_fn fast path): every slot takes SQ triples.
template <class C, int ROLES>
struct QueryMap {
    static constexpr bool kMixed = ROLES == 3 && C::TQP >= 2;
    static constexpr bool kSplit = ROLES == 3 && C::TQP == 1;
    static constexpr int kTriplesPerSlot = kMixed ? C::TQP : C::SQ;
    static constexpr int kRoleSlots = kSplit ? C::NS / 2 : C::NS;
    static constexpr int kTriplesPerGroup = C::TS ? kTriplesPerSlot : kRoleSlots * kTriplesPerSlot;
    __device__ static __forceinline__ bool is_head(int slot, int qi) {
        if (kMixed) return qi < C::TQP;
        if (kSplit) return slot < C::NS / 2;
        return ROLES == 1;
    }
    __device__ static __forceinline__ int triple(int slot, int qi) {      // offset inside the group
        if (kMixed) return (C::TS ? 0 : slot * C::TQP) + (qi % C::TQP);
        if (kSplit) return (slot % (C::NS / 2)) * C::SQ + qi;
        return slot * C::SQ + qi;
    }
};

template <int MODEL, class C, int ROLES>
__global__ void __launch_bounds__((C::CW + 1) * 32, 1) sweep_kernel(const SweepArgs args, const __grid_constant__ CUtensorMap tmap) {
    using QM = QueryMap<C, ROLES>;
    using SM = SweepSmem<MODEL, C>;
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment for the TMA swizzle; plain pointer arithmetic keeps the shared address space (LDS, not LD)
    SM &sm = *reinterpret_cast<SM *>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    BLP_TS(0);

    if (tid == 0) {
        for (int s = 0; s < SM::kFull; ++s) mbar_init(&sm.full_bar[s], 1);
        for (int s = 0; s < SM::ST; ++s)
            mbar_init(&sm.empty_bar[s], C::TS ? C::CW / C::NS : C::CW);   // tile split: one slot's warps consume a tile
        mbar_fence_init();
    }
    __syncthreads();

    griddep_launch_dependents();
    // everything this grid shares with the previous launch of the stream -- the counter workspace, the ticket, the output
    // arrays -- is touched only after griddep_wait(); in the fused step that is the end of a CTA's first segment, so its
    // prologue and scoring overlap the tail of the previous batch.  Inputs that a previous KERNEL may have produced
    // (precomputed true scores) and direct score output wait up front.
    if (!args.fuse_true || args.scores_out) griddep_wait();

    // ---- this CTA's share of the (group, tile) list ---------------------------------------------
    const long long ntiles = (args.n_local + kCT - 1) / kCT;
    const long long total = ntiles * args.groups;
    // contiguous shares, the CTAs with one item more FIRST: CTAs are dispatched in blockIdx order, so with overlapping
    // launches (blp_plan_set_overlap) the last-started CTA of a batch -- the one that decides when the batch completes --
    // is a short one (with the rounding total * b / grid the last CTA always got the long share: no gain from overlap)
    const long long share = total / gridDim.x, extra = total % gridDim.x, bid = blockIdx.x;
    const long long id_begin = bid * share + (bid < extra ? bid : extra), id_end = id_begin + share + (bid < extra ? 1 : 0);

    if (warp == C::CW) {
        // ===================== producer warp =====================
        int it = 0;
        for (long long id = id_begin; id < id_end; ++id, ++it) {
            const long long tile = id % ntiles;
            const int buf = it % SM::ST;
            const uint32_t use = (uint32_t)(it / SM::ST);
            uint64_t *full = &sm.full_bar[it % SM::kFull];
            mbar_wait(&sm.empty_bar[buf], (use & 1u) ^ 1u);
            const long long base = tile * kCT;
            const int rows = (int)min((long long)kCT, args.n_local - base);
            float *dst = &sm.ctile[buf][0];
            if (MODEL == BLP_MODEL_TRANSE && args.use_tma) {
                // four tensor copies of [128 rows][32 floats]; rows past the end of the table arrive as zeros
                if (lane == 0) {
                    mbar_arrive_expect_tx(full, (uint32_t)(kCT * kD * 4));
#pragma unroll
                    for (int cb = 0; cb < 4; ++cb)
                        tma_tensor2d_g2s(dst + cb * (kCT * 32), &tmap, cb * 32, (int)base, full);
                }
            } else {
                // 16-byte chunks; a warp-wide load covers one 512-byte row
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
                        const int row = batch * 16 + u;
                        if (MODEL == BLP_MODEL_TRANSE) {
                            *reinterpret_cast<float4 *>(dst + TileLayout<MODEL>::chunk(row, lane)) = v[u];
                        } else {
                            float *drow = dst + row * kPitch;
                            drow[perm_pos<MODEL>(4 * lane + 0)] = v[u].x;
                            drow[perm_pos<MODEL>(4 * lane + 1)] = v[u].y;
                            drow[perm_pos<MODEL>(4 * lane + 2)] = v[u].z;
                            drow[perm_pos<MODEL>(4 * lane + 3)] = v[u].w;
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(full);
            }
        }
        return;
    }

    // ===================== consumer warps =====================
    // RS warps share a slot and split the 128 tile rows between them
    const int slot = warp / C::RS, rowgrp = warp % C::RS;
    const float *qv = &sm.qv[slot * C::TQP][0][0][0];
    const int row_off = rowgrp * C::TC * 32 + lane;                        // this lane's first row inside a tile
    const long long tail_base = (ROLES == 3) ? args.tail_off : 0;          // output slot of tail query i is tail_off + i

    int it = 0;
    long long id = id_begin;
    BLP_TS(1);
    int nseg = 0;
    while (id < id_end) {
        ++nseg;
        // ---- segment = the run of this CTA's tiles that belong to one triple group
        const long long grp = id / ntiles;
        const long long seg_end = min(id_end, (grp + 1) * ntiles);
        const long long t0 = grp * QM::kTriplesPerGroup;                   // first triple of this group
        const bool seg_first = id == grp * ntiles;                         // this CTA owns the group's first tile: it reports the true scores

        consumer_bar_sync();                      // everyone is done with the previous group's vectors
        // (1) the h / t / r row of every query of this group, resolved once (one round of index loads)
        if (tid < C::NQ * 3) {
            const int ql = tid / 3, which = tid - 3 * ql;
            const int s_ = ql / C::SQ, qi = ql % C::SQ;
            const bool hq = QM::is_head(s_, qi);
            const long long tr = t0 + QM::triple(s_, qi);
            const long long trc = tr < args.b ? tr : args.b - 1;
            const RowRef &X = which == 0 ? (hq ? args.h : args.h2) : which == 1 ? (hq ? args.t : args.t2) : (hq ? args.r : args.r2);
            sm.rowp[ql][which] = X.row(trc, kD);
            sm.rowok[ql][which] = X.in_range(trc) ? 1 : 0;
        }
        consumer_bar_sync();
        // (2) one warp per query row, one 16-byte chunk per lane: the rows are read once; the folded operands go to
        // their place, the per-position terms of the true-triple score are parked for the summation chains.
        // (3) the reference's summation order, the chains of 4 queries side by side in the warp's lanes: lane u runs
        // the 128 dependent adds of torch.norm(p=1) for the round's u-th query; for torch.sum, lanes 8u .. 8u+7 run
        // ATen's 8-lane x 4-accumulator cascade of the u-th query.
        constexpr bool kHalves = MODEL == BLP_MODEL_COMPLEX || MODEL == BLP_MODEL_SIMPLE;
        constexpr int kPerWarp = (C::NQ + C::CW - 1) / C::CW;          // this warp's queries: ql = warp + CW * u
        constexpr int kTermRows = Stages<MODEL, C>::kTermRows;
        constexpr int L = kHalves ? kD / 2 : kD;
        const int tm0 = warp * Stages<MODEL, C>::kRound;                // this warp's term rows
        constexpr int kR = Stages<MODEL, C>::kRound;                    // queries per warp and round (<= 4)
#pragma unroll 1
        for (int u0 = 0; u0 < kPerWarp; u0 += kR) {
#pragma unroll
            for (int uu = 0; uu < kR; ++uu) {
                const int ql = warp + C::CW * (u0 + uu);
                if (u0 + uu >= kPerWarp || ql >= C::NQ) continue;     // warp-uniform
                const int s_ = ql / C::SQ, qi = ql % C::SQ;
                const bool hp = QM::is_head(s_, qi);
                const long long tr = t0 + QM::triple(s_, qi);
                float *qp = &sm.qv[ql >> 1][0][0][ql & 1];
                const int c = kHalves ? (lane & 15) : lane;
                if (tr < args.b) {
                    const float *h = sm.rowp[ql][0], *t = sm.rowp[ql][1], *r = sm.rowp[ql][2];
                    const float4 h0 = __ldg(reinterpret_cast<const float4 *>(h) + c), t0v = __ldg(reinterpret_cast<const float4 *>(t) + c),
                                 r0 = __ldg(reinterpret_cast<const float4 *>(r) + c);
                    float4 h1 = h0, t1 = t0v, r1 = r0;
                    if (kHalves) {
                        h1 = __ldg(reinterpret_cast<const float4 *>(h + 64) + c);
                        t1 = __ldg(reinterpret_cast<const float4 *>(t + 64) + c);
                        r1 = __ldg(reinterpret_cast<const float4 *>(r + 64) + c);
                    }
                    if (!kHalves || lane < 16) {
                        const float ha[4] = {h0.x, h0.y, h0.z, h0.w}, ta[4] = {t0v.x, t0v.y, t0v.z, t0v.w}, ra[4] = {r0.x, r0.y, r0.z, r0.w};
                        const float hb[4] = {h1.x, h1.y, h1.z, h1.w}, tb[4] = {t1.x, t1.y, t1.z, t1.w}, rb[4] = {r1.x, r1.y, r1.z, r1.w};
                        float o[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (hp) fold_vals<MODEL, true>(ha[k], ta[k], ra[k], hb[k], tb[k], rb[k], 4 * c + k, qp);
                            else fold_vals<MODEL, false>(ha[k], ta[k], ra[k], hb[k], tb[k], rb[k], 4 * c + k, qp);
                            o[k] = true_term<MODEL>(ha[k], ta[k], ra[k], hb[k], tb[k], rb[k]);
                        }
                        if (args.fuse_true) *reinterpret_cast<float4 *>(&sm.tm[tm0 + uu][4 * c]) = make_float4(o[0], o[1], o[2], o[3]);
                    }
                } else if (!kHalves || lane < 16) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int j = 4 * c + k;
                        qp[j * 2] = 0.0f;
                        qp[(kD + j) * 2] = 0.0f;
                        if (kHalves) {
                            qp[(64 + j) * 2] = 0.0f;
                            qp[(kD + 64 + j) * 2] = 0.0f;
                        }
                    }
                }
            }
            if (!args.fuse_true) continue;
            __syncwarp();
            float sv = 0.0f;
            int uu_mine = -1;                                          // the chain this lane finishes
            if (MODEL == BLP_MODEL_TRANSE) {
                if (lane < kR) {
                    const float *tu = &sm.tm[tm0 + lane][0];
#pragma unroll 8
                    for (int j = 0; j < kD; j += 4) {
                        const float4 v = *reinterpret_cast<const float4 *>(tu + j);
                        sv = fadd(fadd(fadd(fadd(sv, v.x), v.y), v.z), v.w);
                    }
                    sv = -sv;
                    uu_mine = lane;
                }
            } else {
                const int uu = lane >> 3, l = lane & 7;
                const float *tu = &sm.tm[tm0 + (uu < Stages<MODEL, C>::kRound ? uu : 0)][0];
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < L / 32; ++k)
#pragma unroll
                    for (int a = 0; a < 4; ++a) acc[a] = fadd(acc[a], tu[32 * k + 8 * a + l]);
                const float cl = fadd(fadd(fadd(acc[0], acc[1]), acc[2]), acc[3]);
#pragma unroll
                for (int ll = 0; ll < 8; ++ll) sv = fadd(sv, __shfl_sync(0xffffffffu, cl, (lane & 24) + ll));
                if (MODEL == BLP_MODEL_SIMPLE) sv = fmul(sv, 0.5f);
                if (l == 0 && uu < kR) uu_mine = uu;
            }
            if (uu_mine >= 0) {
                const int ql = warp + C::CW * (u0 + uu_mine);
                if (u0 + uu_mine < kPerWarp && ql < C::NQ) {
                    const int s_ = ql / C::SQ, qi = ql % C::SQ;
                    const long long tr = t0 + QM::triple(s_, qi);
                    // an index outside the table (train.py:137-138 asserts this never happens): NaN compares false
                    // against every candidate, so the query reports gt = ge = 0 and is detectable
                    const float v = (sm.rowok[ql][0] & sm.rowok[ql][1] & sm.rowok[ql][2]) ? sv : __int_as_float(0x7fc00000);
                    sm.st[ql] = tr < args.b ? v : 0.0f;             // written to true_score_out at the end of the segment
                }
            }
            __syncwarp();
        }
        if (!args.fuse_true && tid < C::NQ) {
            const int s_ = tid / C::SQ, qi = tid % C::SQ;
            const long long tr = t0 + QM::triple(s_, qi);
            const long long qo = (QM::is_head(s_, qi) ? 0 : tail_base) + tr;
            sm.st[tid] = (args.true_score && tr < args.b) ? args.true_score[qo] : 0.0f;
        }
        if (nseg == 1) BLP_TS(2);
        consumer_bar_sync();
        if (nseg == 1) BLP_TS(3);

        float st[C::SQ];
        long long qo[C::SQ];                      // output row of each of this thread's queries, -1 = padding
        bool any = false;
#pragma unroll
        for (int q = 0; q < C::SQ; ++q) {
            st[q] = sm.st[slot * C::SQ + q];
            const long long tr = t0 + QM::triple(slot, q);
            qo[q] = tr < args.b ? (QM::is_head(slot, q) ? 0 : tail_base) + tr : -1;
            any |= qo[q] >= 0;
        }
        int cgt[C::SQ], cge[C::SQ];
#pragma unroll
        for (int q = 0; q < C::SQ; ++q) cgt[q] = cge[q] = 0;
        // Mixed role: do the head-prediction triples of EVERY slot of this group share one relation row per slot?  Then the
        // first rounding of head prediction -- fl(candidate + r) for TransE, fl(candidate * r) for the bilinear models --
        // is computed once per slot (score_tile<..., SHARED_R>).  The decision is taken per
        // group, not per slot, so that all warps of the CTA run the same loop nest at any time: the two nests are
        // ~20 KB of code each and evict each other from the instruction cache when both are hot (measured: 456 us
        // instead of 284 / 314 us per 1,024-triple launch).  rank_sweep(sort_by_relation=True) pads every relation's run
        // to a multiple of 4 triples, which makes every group uniform.
#ifndef BLP_SHARED_R
#define BLP_SHARED_R 1
#endif
        bool shared_r = false;
        if (BLP_SHARED_R && QM::kMixed && C::TQP >= 2 && (C::TC == 4 || C::CW > 8)) {   // only the 32-triple groups gain
            shared_r = true;
#pragma unroll
            for (int s_ = 0; s_ < C::NS; ++s_) {
                const long long tr0 = t0 + QM::triple(s_, 0);
                if (tr0 >= args.b) continue;                       // slot past the end of the batch: no work
                bool ok = tr0 + C::TQP <= args.b;
                const float *r0 = sm.rowp[s_ * C::SQ][2];
#pragma unroll
                for (int q = 1; q < C::TQP; ++q) ok &= sm.rowp[s_ * C::SQ + q][2] == r0;
                shared_r &= ok;
            }
        }

        for (; id < seg_end; ++id, ++it) {
            if (C::TS && (it % C::NS) != slot) continue;      // tile split: the other slot's tile
            const long long tile = id % ntiles;
            const int buf = it % SM::ST;
            mbar_wait(&sm.full_bar[it % SM::kFull], (uint32_t)(it / SM::kFull) & 1u);
            if (it == 0) BLP_TS(4);
            float s[C::SQ][C::TC];
            if (any) {                            // warp-uniform: slots past the end of the batch have no queries
                f2 sp[C::TQP][C::TC];
                const TileView<MODEL> tv(&sm.ctile[buf][0], row_off);
#ifndef BLP_FORCE_SHARED
#define BLP_FORCE_SHARED 0      // timing experiment only: the shared-relation loop nest alone (wrong results unless every slot shares r)
#endif
                constexpr bool kCanShare = BLP_SHARED_R && QM::kMixed && C::TQP >= 2 && (C::TC == 4 || C::CW > 8);
                if (kCanShare && (BLP_FORCE_SHARED || shared_r))
                    score_tile<MODEL, kRoleMixed, C::TQP, C::TC, true>(tv, qv, args.negzero2, sp);
                else if (QM::kMixed && !(kCanShare && BLP_FORCE_SHARED)) score_tile<MODEL, kRoleMixed, C::TQP, C::TC>(tv, qv, args.negzero2, sp);
                else if (QM::is_head(slot, 0)) score_tile<MODEL, kRoleHead, C::TQP, C::TC>(tv, qv, args.negzero2, sp);
                else score_tile<MODEL, kRoleTail, C::TQP, C::TC>(tv, qv, args.negzero2, sp);
#pragma unroll
                for (int q = 0; q < C::TQP; ++q)
#pragma unroll
                    for (int i = 0; i < C::TC; ++i) {
                        unpack2(sp[q][i], s[2 * q][i], s[2 * q + 1][i]);
                        if (MODEL == BLP_MODEL_TRANSE) {              // score = -||.||_1
                            s[2 * q][i] = -s[2 * q][i];
                            s[2 * q + 1][i] = -s[2 * q + 1][i];
                        }
                    }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty_bar[buf]);
            if (!any) continue;

            const long long base = tile * kCT + row_off;
            if (args.scores_out) {
#pragma unroll
                for (int q = 0; q < C::SQ; ++q) {
                    if (qo[q] >= 0) {
#pragma unroll
                        for (int i = 0; i < C::TC; ++i) {
                            const long long cand = base + 32 * i;
                            if (cand < args.n_local) args.scores_out[qo[q] * args.ld_scores + cand] = s[q][i];
                        }
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < C::TC; ++i) {
                    const bool valid = base + 32 * i < args.n_local;
#pragma unroll
                    for (int q = 0; q < C::SQ; ++q) {
                        cgt[q] += (valid && s[q][i] > st[q]) ? 1 : 0;
                        cge[q] += (valid && s[q][i] >= st[q]) ? 1 : 0;
                    }
                }
            }
        }
        griddep_wait();                           // the previous launch of the stream is complete: outputs / workspace are ours
        if (args.fuse_true && seg_first && args.true_score_out && tid < C::NQ) {
            const int s_ = tid / C::SQ, qi = tid % C::SQ;
            const long long tr = t0 + QM::triple(s_, qi);
            if (tr < args.b) args.true_score_out[(QM::is_head(s_, qi) ? 0 : tail_base) + tr] = sm.st[tid];
        }
        if (!args.scores_out) {
#pragma unroll
            for (int q = 0; q < C::SQ; ++q) {
                const int a = __reduce_add_sync(0xffffffffu, cgt[q]);
                const int c = __reduce_add_sync(0xffffffffu, cge[q]);
                if (lane == 0 && qo[q] >= 0 && (a | c)) {
                    atomicAdd(&args.gt[qo[q]], a);
                    atomicAdd(&args.ge[qo[q]], c);
                }
            }
        }
    }

    BLP_TS(5);
    if (args.dbg && tid == 0) { args.dbg[(size_t)blockIdx.x * 16 + 8] = (unsigned long long)nseg; args.dbg[(size_t)blockIdx.x * 16 + 9] = (unsigned long long)(id_end - id_begin); }
    if (!args.fuse_epilogue) return;
    // ---- fused epilogue: the last CTA to finish turns the accumulated counters into the step's results
    // (utils.py:103-109, train.py:154-157) and leaves the workspace zeroed for the next call
    // scratch in the (now idle) query-operand / true-score areas of the dynamic shared memory
    volatile int *s_last = reinterpret_cast<volatile int *>(&sm.st[0]);
    double (*s_red)[9] = reinterpret_cast<double (*)[9]>(&sm.qv[0][0][0][0]);
    consumer_bar_sync();                      // this CTA's counter updates are issued (CTA-scope order) ...
    if (tid == 0) {
        __threadfence();                      // ... and made visible device-wide before the ticket (cumulative fence)
        const unsigned int done = atomicAdd(args.ticket, 1u);
        *s_last = (done == gridDim.x - 1) ? 1 : 0;
    }
    consumer_bar_sync();
    BLP_TS(6);
    if (!*s_last) return;
    __threadfence();
    double acc[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) acc[j] = 0.0;
    for (long long i = tid; i < args.out_len; i += C::CW * 32) {
        const bool live = i < args.b || i >= args.tail_off;       // [b, tail_off) is a gap when the outputs are slices
        if (!live) continue;
        const int g = __ldcg(args.gt + i), e = __ldcg(args.ge + i);
        args.gt[i] = 0;
        args.ge[i] = 0;
        args.gt_out[i] = g;
        args.ge_out[i] = e;
        const float avg = fmul((float)((long long)g + 1 + (long long)e), 0.5f);
        const float rr = __frcp_rn(avg);
        if (args.recip) args.recip[i] = rr;
        acc[0] += (double)rr;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < args.kv.nk) {
                const bool hit = avg <= (float)args.kv.k[j];
                if (hit) acc[1 + j] += 1.0;
                if (args.hits) args.hits[i * args.kv.nk + j] = hit;
            }
    }
    if (args.sums) {
#pragma unroll
        for (int j = 0; j < 9; ++j) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
            if (lane == 0) s_red[warp][j] = acc[j];
        }
        consumer_bar_sync();
        if (tid <= args.kv.nk) {
            double tot = 0.0;
            for (int w = 0; w < C::CW; ++w) tot += s_red[w][tid];
            args.sums[tid] = tot;
        }
    }
    if (tid == 0) *args.ticket = 0u;
    BLP_TS(7);
}

// ---- host side ----------------------------------------------------------------
unsigned long long *debug_timestamp_buffer();
static unsigned long long *tl_dbg_get() { return debug_timestamp_buffer(); }
// Per-launch host work is kept to the launch itself (it is a large share of a 20 us eval batch): the SM count and
// the shared-memory opt-in are cached per device, the tensor map per (table pointer, rows) in a small thread-local
// cache, the tuning environment variables are read once.
constexpr int kMaxDevices = 64;

static int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}

static int num_sms() {
    static std::atomic<int> cached[kMaxDevices];
    const int dev = current_device();
    int n = cached[dev].load(std::memory_order_relaxed);
    if (n <= 0) {
        n = 148;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
        cached[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (libcuda is not linked)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// 2-D view of the table shard: [n_local rows][128 floats], boxes of [128 rows][32 floats], 128-byte swizzle.
// The descriptor depends only on (pointer, rows), so a handful of them are remembered per host thread.
static bool make_table_tmap(CUtensorMap *tm, const float *ent, long long n_local) {
    struct Entry { const float *ent; long long n; CUtensorMap map; };
    constexpr int kSlots = 8;
    static thread_local Entry cache[kSlots];
    static thread_local int next = 0;
    for (int i = 0; i < kSlots; ++i)
        if (cache[i].ent == ent && cache[i].n == n_local) { *tm = cache[i].map; return true; }
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)kD, (cuuint64_t)n_local};
    const cuuint64_t strides[1] = {(cuuint64_t)kD * 4};
    const cuuint32_t box[2] = {32, (cuuint32_t)kCT};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ent), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    cache[next] = Entry{ent, n_local, *tm};
    next = (next + 1) % kSlots;
    return true;
}

template <int MODEL, class C, int ROLES>
static int launch_sweep_cfg(const SweepArgs &a, cudaStream_t st) {
    const size_t smem = sizeof(SweepSmem<MODEL, C>) + 1024;
    static std::atomic<unsigned long long> attr_done{0};        // bit per device: opt-in to > 48 KB dynamic smem
    const int dev = current_device();
    if (!((attr_done.load(std::memory_order_relaxed) >> dev) & 1ull)) {
        BLP_CUDA(cudaFuncSetAttribute(sweep_kernel<MODEL, C, ROLES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done.fetch_or(1ull << dev, std::memory_order_relaxed);
    }
    SweepArgs args = a;
    args.dbg = tl_dbg_get();
    const int tg = QueryMap<C, ROLES>::kTriplesPerGroup;
    args.groups = (a.b + tg - 1) / tg;
    const long long ntiles = (a.n_local + kCT - 1) / kCT;
    const long long items = ntiles * args.groups;
    if (items == 0) return BLP_OK;
    const long long sms = num_sms();
    const unsigned grid = (unsigned)(items < sms ? items : sms);
    CUtensorMap tmap;
    if (MODEL == BLP_MODEL_TRANSE && args.use_tma) {
        if (!make_table_tmap(&tmap, a.ent, a.n_local)) args.use_tma = 0;
    }
    if (!(MODEL == BLP_MODEL_TRANSE && args.use_tma)) memset(&tmap, 0, sizeof(tmap));
    prof_begin(1, st);
    if (args.overlap) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3((C::CW + 1) * 32);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        BLP_CUDA(cudaLaunchKernelEx(&cfg, sweep_kernel<MODEL, C, ROLES>, args, tmap));
    } else {
        sweep_kernel<MODEL, C, ROLES><<<grid, (C::CW + 1) * 32, smem, st>>>(args, tmap);
    }
    prof_end(1, st);
    count_launch();
    BLP_CUDA(cudaGetLastError());
    return BLP_OK;
}

static int env_sweep_cfg() {
    static const int v = []() {
        const char *e = getenv("BLP_SWEEP_CFG");       // tuning aid: force one register tile
        return (e && e[0] >= '0' && e[0] <= '6') ? e[0] - '0' : -1;
    }();
    return v;
}

// Register-tile shape by batch size: the smallest slot that holds the batch without padding queries.
// a.force_cfg (or BLP_SWEEP_CFG=0..6) forces one: 0..4 = 2 / 4 / 8 / 16 / 32 triples per table pass, 5 / 6 = the
// tile-split forms of 2 / 4.
template <int MODEL>
static int launch_sweep(const SweepArgs &a, cudaStream_t st) {
    if (a.roles == 1) return launch_sweep_cfg<MODEL, Cfg<4, 4>, 1>(a, st);   // single role: full tiles only
    if (a.roles == 2) return launch_sweep_cfg<MODEL, Cfg<4, 4>, 2>(a, st);
    // up to 256 triples the 16-triple groups of Cfg<4,2> give twice as many (group, tile) work items, which evens out
    // the per-CTA shares (E = 64: 44 vs 50 us, E = 256: 110 vs 124 us); beyond that the larger register tile of
    // Cfg<4,4> and its shared-relation path win (E = 1024 in relation order: 0.321 vs 0.390 ms)
    // 2 / 4 triples per pass (HBM-bound shapes): the tile-split forms -- every thread holds the head AND tail pairs of
    // the pass for one row (2 / 4 independent chains, one candidate load for all of them), the two slots of a CTA
    // alternate on the tiles: 4.8 M rows at pass size 2: 0.409 -> 0.341 ms per pass (7.2 TB/s), pass size 4: 0.575 ->
    // 0.509 ms (profiles/r02_sweep_warp_variants.txt)
    int cfg = a.b <= 2 ? 5 : a.b <= 4 ? 6 : a.b <= 8 ? 2 : a.b <= 256 ? 3 : 4;
    if (a.force_cfg >= 0 && a.force_cfg <= 6) cfg = a.force_cfg;
    else if (env_sweep_cfg() >= 0) cfg = env_sweep_cfg();
    switch (cfg) {
    case 0: return launch_sweep_cfg<MODEL, Cfg<1, 1>, 3>(a, st);             //  2 triples / group (split roles)
    case 5: return launch_sweep_cfg<MODEL, Cfg<2, 1, kCW, true>, 3>(a, st);   //  2 triples / group (tile split: head + tail pair per thread)
    case 6: return launch_sweep_cfg<MODEL, Cfg<4, 1, kCW, true>, 3>(a, st);   //  4 triples / group (tile split: 2 head + 2 tail pairs per thread)
    case 1: return launch_sweep_cfg<MODEL, Cfg<2, 1>, 3>(a, st);             //  4
    case 2: return launch_sweep_cfg<MODEL, Cfg<4, 1>, 3>(a, st);             //  8
    case 3: return launch_sweep_cfg<MODEL, Cfg<4, 2>, 3>(a, st);             // 16
    default: return launch_sweep_cfg<MODEL, Cfg<4, 4>, 3>(a, st);            // 32
    }
}

int launch_sweep_dyn(int model, const SweepArgs &a, cudaStream_t st) {
    switch (model) {
    case BLP_MODEL_TRANSE: return launch_sweep<BLP_MODEL_TRANSE>(a, st);
    case BLP_MODEL_DISTMULT: return launch_sweep<BLP_MODEL_DISTMULT>(a, st);
    case BLP_MODEL_COMPLEX: return launch_sweep<BLP_MODEL_COMPLEX>(a, st);
    default: return launch_sweep<BLP_MODEL_SIMPLE>(a, st);
    }
}

// Cfg index whose group holds `group_triples` triples (the reference's eval batch as a table-pass size), -1 = auto
int sweep_cfg_for_group(long long group_triples) {
    if (group_triples <= 0) return -1;
    return group_triples <= 2 ? 5 : group_triples <= 4 ? 6 : group_triples <= 8 ? 2 : group_triples <= 16 ? 3 : 4;
}

static thread_local unsigned long long *tl_dbg = nullptr;
unsigned long long *debug_timestamp_buffer() { return tl_dbg; }
void set_debug_timestamp_buffer(unsigned long long *p) { tl_dbg = p; }

int sweep_env_use_tma() {
    static const int v = []() {
        const char *e = getenv("BLP_EVAL_PRODUCER");
        return (e && (e[0] == 'l' || e[0] == 'L')) ? 0 : 1;   // "ldg"
    }();
    return v;
}

}  // namespace blp
