// blp_fast.cu -- tensor-core ("fast") mode of the full-entity sweep for the bilinear models
// (north_star: "tensor cores used only for the DistMult/ComplEx case
// reformulated as a dense contraction").
//
// DistMult / ComplEx / SimplE scores are linear in the candidate row once the query side is folded:
//     score(q, e) = sum_d c_q[d] * e[d]           (models.py:226-248 with the candidate factored out)
// so a sweep is S = C (Q x 128) . E^T (128 x N) followed by the same rank counting as the exact mode.
// Folding changes the fp32 roundings of the reference ((h*r)*t vs (r*t)*h, models.py:227), so this mode is
// NOT bit-exact by construction; it is opt-in and its parity is tolerance-classified:
// |score_fast - score_ref| <= 1e-5 * sum|terms|, rank differences only inside that band.
//
// Arithmetic: split-FP16 ("3xFP16").  Every fp32 operand x is scaled by a power of two s (exact) and split as
// s*x = hi + lo with hi = fp16(s*x), lo = fp16(s*x - hi): 22 significant bits, the same as the classic 3xTF32
// split, but kind::f16 runs at twice the kind::tf32 rate and the operands take half the bytes.
// hi*hi + lo*hi + hi*lo is accumulated in fp32 in TMEM (the dropped lo*lo term is ~2^-22 relative).  The scales
// keep fp16 in range for any input magnitude: one per entity table (max |e| -> 2^14, blp_fast_prepare_table) and
// one per query row (max |c_q| -> 2^14, fold kernel); the epilogue compares the raw accumulator against the true
// score pre-multiplied by (s_table * s_query), so no per-score rescaling is needed.
//
// Kernel (sm_100a, one persistent CTA per SM, 384 threads).  A work item is (256 queries) x (128 candidates):
//   warp 0 lane 0   TMA producer: the query block (2 halves x [hi | lo], 128 KB, resident while the CTA stays on
//                   the same 256 queries) and a 3-stage ring of candidate k-blocks ([hi | lo] boxes of 128 rows x
//                   64 halves, 128-byte swizzle)
//   warp 1 lane 0   tcgen05.mma issuer (kind::f16, M = 128, N = 128, K = 16): every candidate k-block is used by
//                   BOTH 128-query halves (two accumulators), which halves the L2 -> SMEM bytes per flop -- at one
//                   half per CTA the kernel is bound by L2 bandwidth, not by the tensor pipe; accumulators are
//                   double-buffered in TMEM (2 buffers x 2 halves x 128 columns = all 512 columns)
//   warp 2          TMEM allocation
//   warps 4-11      epilogue: tcgen05.ld 32 columns at a time -- TMEM lane = query row, so each thread owns one
//                   query of each half; the two warps of a lane quarter split the column chunks -- compare against
//                   the scaled true score, count, one atomicAdd pair per query and query-block run.  No score
//                   ever leaves the SM.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include "blp_sweep.h"

namespace blp {

#ifndef BLP_FAST_N
#define BLP_FAST_N 128
#endif
constexpr int kFM = 128, kFN = BLP_FAST_N; // rows per MMA (TMEM lanes), candidates per work item
constexpr int kFH = 2;                     // 128-query halves per work item
constexpr int kFBufs = 512 / (kFH * kFN);  // accumulator buffers in TMEM (all 512 columns)
constexpr int kFStages = 3 * 128 / kFN;    // candidate k-block ring (96 KB)
constexpr int kFKB = 64;                   // halves per k-block (one 128-byte swizzle span)
constexpr int kFBox = 128 * kFKB;          // halves in one query TMA box (128 rows x 64 halves = 16 KB)
constexpr int kFBoxB = kFN * kFKB;         // halves in one candidate TMA box (kFN rows x 64 halves)
constexpr int kFEpiWarps = 8;                // epilogue warps: two per TMEM lane quarter, they split the column chunks
constexpr int kFThreads = (4 + kFEpiWarps) * 32;
constexpr int kTmemCols = 512;             // kFBufs buffers x 2 halves x kFN columns
constexpr float kFTargetExp = 14.0f;       // operands are scaled so that max |x| <= 2^14 (fp16 max is 65504)

// PAIR = a cluster of two CTAs running tcgen05.mma.cta_group::2 (M = 256: 128 query rows in each CTA's TMEM): every
// CTA stages only HALF of each candidate k-block (64 of the 128 rows), so the shared-memory operand fetch of an MMA
// drops from 8 KB to 6 KB per SM and the L2 -> SMEM candidate stream is halved; the ring gets twice the stages.
template <bool PAIR>
struct FastCfg {
    static constexpr int kRowsB = PAIR ? kFN / 2 : kFN;      // candidate rows of a k-block staged by one CTA
    static constexpr int kBoxB = kRowsB * kFKB;              // halves in one candidate TMA box
    static constexpr int kStages = PAIR ? 2 * kFStages : kFStages;
    static constexpr int kQ = (PAIR ? 2 : 1) * kFH * kFM;    // queries per work item (256, or 512 over the pair)
};
template <bool PAIR>
struct __align__(1024) FastSmemT {
    static constexpr int kStages = FastCfg<PAIR>::kStages;
    __half a[kFH][2][2][kFBox];            // [query half][hi | lo][k-block] query block (this CTA's 256 queries)
    __half b[kStages][2][FastCfg<PAIR>::kBoxB];   // [stage][hi | lo] candidate k-block (this CTA's rows of it)
    uint64_t a_full, a_empty;
    uint64_t b_full[kStages], b_empty[kStages];
    uint64_t d_full[kFBufs], d_empty[kFBufs];
    uint32_t tmem_base;
};

// Worklist of the refine pass: (query row q in [0, 2b), local candidate row) pairs.
struct RefineList {
    unsigned int count;                    // appended entries (may exceed the capacity: the excess is dropped and flagged)
    unsigned int overflow;                 // sticky: set by the refine kernel when count > capacity
    unsigned int pad[2];
    int2 entries[1];
};

struct FastArgs {
    long long n_local, ent_offset, n_pad;  // candidates in this shard, global id of row 0, rows per half of the split table
    long long b, tail_off, q_pad;          // triples, output slot offset of tail queries, rows per half of the split queries
    long long m_tiles, n_tiles;
    const float *true_score;               // indexed by output slot
    const long long *self_id;              // [q_pad] global candidate id of the query's true entity (-1 = padding)
    const float *qscale;                   // [q_pad] s_table * s_query of every query row (powers of two)
    int *gt, *ge;
    float *scores_out;                     // optional (2b, ld_scores) matrix of the fast scores (verification aid)
    long long ld_scores;
    int debug;                             // timing experiments (BLP_FAST_DEBUG): 1 = no epilogue work, 2 = no B loads, 4 = no MMAs
    // ---- exact ranks on the tensor path (filter + refine): candidates whose fast score lies within `band[q]` of the
    // scaled true score are not counted here but appended to the worklist and re-scored in the reference's fp32 order
    int pair;                              // launched as CTA pairs (cta_group::2)
    const float *band;                     // [q_pad] a-priori bound on |fast - reference| in the scaled domain (NULL = plain fast mode)
    RefineList *refine;                    // worklist (header + entries)
    long long refine_cap;
};


// ---- PTX wrappers (tcgen05) ---------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]^T, fp16 inputs, fp32 accumulation
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// ---- cta_group::2 (CTA pair) variants ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t *slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem of both CTAs] (+)= A (128 rows from each CTA) . B (N / 2 rows from each CTA)^T; issued by the leader CTA only
__device__ __forceinline__ void umma_f16_2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs once the pair's tcgen05 operations issued so far have completed
__device__ __forceinline__ void umma_commit2(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((unsigned short)3) : "memory");
}
// shared::cluster address of `bar` in the leader CTA (rank 0 of the pair): the peer bit of the window is bit 24
__device__ __forceinline__ uint32_t leader_bar(const uint64_t *bar) { return smem_u32(bar) & 0xFEFFFFFFu; }
// TMA tile load whose completion bytes are credited to the LEADER's mbarrier (both CTAs of the pair issue it)
__device__ __forceinline__ void tma_tensor2d_g2s_pair(void *dst_smem, const void *tmap, int c0, int c1, const uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(leader_bar(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(const uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(leader_bar(bar)) : "memory");
}

// one lane of the (converged) warp; the same lane every time, so a commit follows the MMAs of its own thread
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
// mbarrier arrive once every tcgen05 operation issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile of [rows][64 halves] written by TMA with the 128-byte swizzle: 8-row groups are
// 1024 bytes apart (SBO), the leading-dimension offset is unused for swizzled K-major layouts.
__device__ __forceinline__ uint64_t umma_smem_desc(const void *p) {
    uint64_t d = (uint64_t)((smem_u32(p) & 0x3FFFFu) >> 4);   // start address, 16-byte units
    d |= (uint64_t)1 << 16;                                    // leading byte offset (ignored)
    d |= (uint64_t)(1024 >> 4) << 32;                          // stride byte offset
    d |= (uint64_t)1 << 46;                                    // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                                    // SWIZZLE_128B
    return d;
}
// kind::f16 with fp16 operands (format 0), fp32 accumulator (format 1), both operands K-major, M x N
__host__ __device__ constexpr uint32_t umma_idesc_f16(int m, int n) {
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// power-of-two scale that maps max_abs to at most 2^14 (1 when the operand is all zeros / not finite)
__device__ __forceinline__ float pow2_scale(float max_abs) {
    if (!(max_abs > 0.0f) || !(max_abs < 3.0e38f)) return 1.0f;
    // max_abs = m * 2^e, m in [0.5, 1): for normal numbers whose scale is a normal number too, e and 2^(target - e) come
    // straight from the exponent field (frexpf / ldexpf are ~100 instructions each and ran in every lane of the fold)
    const unsigned int ef = __float_as_uint(max_abs) >> 23;               // sign bit is 0 here
    if (ef >= 14u && ef <= 254u) return __uint_as_float((267u + (unsigned int)((int)kFTargetExp - 14) - ef) << 23);   // e = ef - 126
    int e;
    frexpf(max_abs, &e);
    return ldexpf(1.0f, (int)kFTargetExp - e);
}
// s*x = hi + lo, both fp16 (round to nearest even); s is a power of two, so s*x is exact
__device__ __forceinline__ void split_f16(float x, float s, __half &hi, __half &lo) {
    const float v = fmul(x, s);
    hi = __float2half_rn(v);
    lo = __float2half_rn(fsub(v, __half2float(hi)));
}

// v[idx] for a runtime idx without spilling v to local memory: a 5-level select tree on the index bits
__device__ __forceinline__ uint32_t select32(const uint32_t (&v)[32], int idx) {
    uint32_t a[16], b[8], c[4], d[2];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = (idx & 16) ? v[i + 16] : v[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = (idx & 8) ? a[i + 8] : a[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] = (idx & 4) ? b[i + 4] : b[i];
#pragma unroll
    for (int i = 0; i < 2; ++i) d[i] = (idx & 2) ? c[i + 2] : c[i];
    return (idx & 1) ? d[1] : d[0];
}

// ---- operand preparation --------------------------------------------------------------------------
// Table workspace: [n_pad rows x 128] fp16 hi, [n_pad x 128] fp16 lo, then a 256-byte header {float scale;
// uint max_abs_bits}.  Rows >= n_local are zero.
struct FastTableHeader {
    float scale;
    unsigned int max_bits;
    unsigned int max_norm_bits;            // max over rows of ||e||_2 (fp32, rounded up a little): the refine band uses it
};

// max_e ||e||_2 over the table rows, one warp per row
__global__ void table_maxnorm_kernel(const float4 *__restrict__ ent, long long n_local, unsigned int *__restrict__ max_norm_bits) {
    const int lane = threadIdx.x & 31;
    float m = 0.0f;
    for (long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n_local;
         row += ((long long)gridDim.x * blockDim.x) >> 5) {
        const float4 v = __ldg(ent + row * (kD / 4) + lane);
        float q = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        m = fmaxf(m, q);
    }
    // sqrt and the fp32 summation above carry a few ulps of error: 1 + 2^-16 covers them with a wide margin
    m = sqrtf(m) * 1.0000153f;
    if (lane == 0 && m > 0.0f && m < 3.0e38f) atomicMax(max_norm_bits, __float_as_uint(m));
}

__global__ void table_maxabs_kernel(const float4 *__restrict__ ent, long long total4, unsigned int *__restrict__ max_bits) {
    float m = 0.0f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(ent + i);
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(max_bits, __float_as_uint(m));   // non-negative floats order like uints
}

__global__ void split_table_kernel(const float4 *__restrict__ ent, long long n_local, long long n_pad,
                                   __half *__restrict__ out, FastTableHeader *__restrict__ hdr) {
    const float s = pow2_scale(__uint_as_float(hdr->max_bits));
    if (blockIdx.x == 0 && threadIdx.x == 0) hdr->scale = s;
    const long long total = n_pad * (kD / 4);
    uint2 *hi_out = reinterpret_cast<uint2 *>(out), *lo_out = reinterpret_cast<uint2 *>(out + n_pad * kD);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / (kD / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < n_local) v = __ldg(ent + i);
        __half h[4], l[4];
        split_f16(v.x, s, h[0], l[0]);
        split_f16(v.y, s, h[1], l[1]);
        split_f16(v.z, s, h[2], l[2]);
        split_f16(v.w, s, h[3], l[3]);
        uint2 ph, pl;
        ph.x = (uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16);
        ph.y = (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16);
        pl.x = (uint32_t)__half_as_ushort(l[0]) | ((uint32_t)__half_as_ushort(l[1]) << 16);
        pl.y = (uint32_t)__half_as_ushort(l[2]) | ((uint32_t)__half_as_ushort(l[3]) << 16);
        hi_out[i] = ph;
        lo_out[i] = pl;
    }
}

// Coefficient of candidate element j for query (triple i, role): the score with the candidate factored out.
// Value form: xk = x[k], xLk = x[L + k] with k = j mod L, `first` = j < L (DistMult: xk = x[j], the rest unused).
template <int MODEL>
__device__ __forceinline__ float fold_coeff_v(bool head_pred, bool first, float hk, float hLk, float tk, float tLk, float rk,
                                              float rLk) {
    if (MODEL == BLP_MODEL_DISTMULT) return head_pred ? fmul(rk, tk) : fmul(hk, rk);   // models.py:227
    if (MODEL == BLP_MODEL_COMPLEX) {   // models.py:230-239; r = (rr | ri), h = (hr | hi), t = (tr | ti)
        const float rr = rk, ri = rLk;
        if (head_pred) {                // candidate = h:  hr (rr tr + ri ti) + hi (rr ti - ri tr)
            const float tr = tk, ti = tLk;
            return first ? fadd(fmul(rr, tr), fmul(ri, ti)) : fsub(fmul(rr, ti), fmul(ri, tr));
        }
        const float hr = hk, hi = hLk;   // candidate = t:  tr (rr hr - ri hi) + ti (rr hi + ri hr)
        return first ? fsub(fmul(rr, hr), fmul(ri, hi)) : fadd(fmul(rr, hi), fmul(ri, hr));
    }
    // SIMPLE, models.py:242-248: 1/2 (hh ra tt + th rb ht); h = (hh | ht), t = (th | tt), r = (ra | rb)
    if (head_pred) return first ? fmul(0.5f, fmul(rk, tLk)) : fmul(0.5f, fmul(tk, rLk));
    return first ? fmul(0.5f, fmul(rLk, hLk)) : fmul(0.5f, fmul(hk, rk));
}
template <int MODEL>
__device__ __forceinline__ float fold_coeff(bool head_pred, const float *__restrict__ h, const float *__restrict__ t,
                                            const float *__restrict__ r, int j) {
    constexpr int L = kD / 2;
    if (MODEL == BLP_MODEL_DISTMULT) return fold_coeff_v<MODEL>(head_pred, true, h[j], 0.f, t[j], 0.f, r[j], 0.f);
    const int k = j & (L - 1);
    return fold_coeff_v<MODEL>(head_pred, j < L, h[k], h[L + k], t[k], t[L + k], r[k], r[L + k]);
}

// sum_j fold_abs(j) * |e[j]| bounds the sum of |terms| the reference adds up for candidate e: the magnitude the
// a-priori error bound of the refine band is relative to (ComplEx folds two products into one coefficient, which may
// cancel; everything else is a single product, so the bound is |coefficient|).
template <int MODEL>
__device__ __forceinline__ float fold_abs_v(bool first, float xk, float xLk, float rk, float rLk, float c) {
    if (MODEL != BLP_MODEL_COMPLEX) return fabsf(c);
    const float rr = fabsf(rk), ri = fabsf(rLk), xr = fabsf(xk), xi = fabsf(xLk);   // x = t (head prediction) or h
    return first ? rr * xr + ri * xi : rr * xi + ri * xr;
}
template <int MODEL>
__device__ __forceinline__ float fold_abs(bool head_pred, const float *__restrict__ h, const float *__restrict__ t,
                                          const float *__restrict__ r, int j, float c) {
    if (MODEL != BLP_MODEL_COMPLEX) return fabsf(c);
    constexpr int L = kD / 2;
    const int k = j & (L - 1);
    const float *x = head_pred ? t : h;
    return fold_abs_v<MODEL>(j < L, x[k], x[L + k], r[k], r[L + k], c);
}

// One CTA of 128 threads per query row: rows [0, b) predict heads, [b, 2b) predict tails, the rest is padding.
// With `true_score` given, the block of head query i also produces the exact true-triple score of triple i (warp 0,
// true_score_warp128) and resets the counters of both of its queries, which saves the separate true-score launch.
template <int MODEL>
__global__ void __launch_bounds__(kD) fold_queries_kernel(const RowRef hr, const RowRef tr, const RowRef rr,
                                                          const long long *__restrict__ triples, long long b,
                                                          long long q_pad, __half *__restrict__ qsplit,
                                                          long long *__restrict__ self_id, float *__restrict__ qscale,
                                                          const FastTableHeader *__restrict__ table_hdr,
                                                          long long tail_off, float *__restrict__ true_score,
                                                          int *__restrict__ gt, int *__restrict__ ge,
                                                          float *__restrict__ band, float kappa) {
    __shared__ float s_max[kD / 32];
    __shared__ float s_sq[kD / 32];
    __shared__ __align__(16) float s_terms[kD];
    const long long q = blockIdx.x;
    const int j = threadIdx.x;
    float c = 0.0f, ca = 0.0f;
    long long self = -1;
    if (q < 2 * b) {
        const bool head_pred = q < b;
        const long long i = head_pred ? q : q - b;
        const float *h = hr.row(i, kD), *t = tr.row(i, kD), *r = rr.row(i, kD);
        c = fold_coeff<MODEL>(head_pred, h, t, r, j);
        ca = fold_abs<MODEL>(head_pred, h, t, r, j, c);
        self = triples[i * 3 + (head_pred ? 0 : 1)];
        if (true_score && head_pred && j < 32) {
            float s = true_score_warp128<MODEL>(h, t, r, s_terms, j);
            if (j == 0) {
                if (!(hr.in_range(i) && tr.in_range(i) && rr.in_range(i))) s = __int_as_float(0x7fc00000);
                true_score[i] = s;
                true_score[tail_off + i] = s;
                gt[i] = 0; gt[tail_off + i] = 0;
                ge[i] = 0; ge[tail_off + i] = 0;
            }
        }
    }
    // per-row power-of-two scale from the row's max |c|
    float m = fabsf(c), n2 = ca * ca;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    }
    if ((j & 31) == 0) { s_max[j >> 5] = m; s_sq[j >> 5] = n2; }
    __syncthreads();
    m = fmaxf(fmaxf(s_max[0], s_max[1]), fmaxf(s_max[2], s_max[3]));
    const float sq = pow2_scale(m);
    __half hi, lo;
    split_f16(c, sq, hi, lo);
    qsplit[q * kD + j] = hi;
    qsplit[(q_pad + q) * kD + j] = lo;
    if (j == 0) {
        self_id[q] = self;
        qscale[q] = fmul(sq, table_hdr->scale);
        // refine band, scaled like the accumulator: kappa * ||fold_abs||_2 * max_e ||e||_2 >= kappa * sum|terms| for every
        // candidate (Cauchy-Schwarz); 1.001 covers the fp32 roundings of the norm itself
        const float cn = sqrtf((s_sq[0] + s_sq[1]) + (s_sq[2] + s_sq[3])) * 1.001f;
        band[q] = kappa * (cn * sq) * (__uint_as_float(table_hdr->max_norm_bits) * table_hdr->scale);
    }
}

// The same folding with one WARP per triple (rows 16-byte aligned): the h / t / r rows are read once for both
// queries of the triple (one float4 per lane and row; the halves models fetch the partner half by shuffle), the row
// maximum and the band norm are warp reductions (no CTA barrier), hi / lo go out as 8-byte stores.  Per element the
// operations are fold_coeff_v / split_f16 as above, so the operands are bit-identical to fold_queries_kernel's.
// Jobs [0, b) are triples; jobs [b, b + q_pad - 2b) are the zero padding rows of the query table.
constexpr int kFoldWarps = 8;
__device__ __forceinline__ float4 shfl_xor4(const float4 v, int m) {
    return make_float4(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m),
                       __shfl_xor_sync(0xffffffffu, v.z, m), __shfl_xor_sync(0xffffffffu, v.w, m));
}
__device__ __forceinline__ void emit_folded_row(const float (&c)[4], const float (&ca)[4], long long q, long long self,
                                                long long q_pad, __half *__restrict__ qsplit, long long *__restrict__ self_id,
                                                float *__restrict__ qscale, float *__restrict__ band, float kappa,
                                                float table_scale, float table_norm, int lane) {
    float m = fmaxf(fmaxf(fabsf(c[0]), fabsf(c[1])), fmaxf(fabsf(c[2]), fabsf(c[3])));
    float n2 = (ca[0] * ca[0] + ca[1] * ca[1]) + (ca[2] * ca[2] + ca[3] * ca[3]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    }
    const float sq = pow2_scale(m);
    __half hi[4], lo[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) split_f16(c[u], sq, hi[u], lo[u]);
    uint2 ph, pl;
    ph.x = (uint32_t)__half_as_ushort(hi[0]) | ((uint32_t)__half_as_ushort(hi[1]) << 16);
    ph.y = (uint32_t)__half_as_ushort(hi[2]) | ((uint32_t)__half_as_ushort(hi[3]) << 16);
    pl.x = (uint32_t)__half_as_ushort(lo[0]) | ((uint32_t)__half_as_ushort(lo[1]) << 16);
    pl.y = (uint32_t)__half_as_ushort(lo[2]) | ((uint32_t)__half_as_ushort(lo[3]) << 16);
    reinterpret_cast<uint2 *>(qsplit + q * kD)[lane] = ph;
    reinterpret_cast<uint2 *>(qsplit + (q_pad + q) * kD)[lane] = pl;
    if (lane == 0) {
        self_id[q] = self;
        qscale[q] = fmul(sq, table_scale);
        // refine band as in fold_queries_kernel (Cauchy-Schwarz bound; 1.001 covers the fp32 roundings of the norm)
        band[q] = kappa * ((sqrtf(n2) * 1.001f) * sq) * table_norm;
    }
}

template <int MODEL>
__global__ void __launch_bounds__(kFoldWarps * 32) fold_triples_kernel(const RowRef hr, const RowRef tr, const RowRef rr,
                                                                       const long long *__restrict__ triples, long long b,
                                                                       long long q_pad, __half *__restrict__ qsplit,
                                                                       long long *__restrict__ self_id, float *__restrict__ qscale,
                                                                       const FastTableHeader *__restrict__ table_hdr,
                                                                       long long tail_off, float *__restrict__ true_score,
                                                                       int *__restrict__ gt, int *__restrict__ ge,
                                                                       float *__restrict__ band, float kappa) {
    __shared__ __align__(16) float s_terms[kFoldWarps][kD];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long job = (long long)blockIdx.x * kFoldWarps + warp;
    if (job >= b + (q_pad - 2 * b)) return;
    const float table_scale = table_hdr->scale, table_norm = __uint_as_float(table_hdr->max_norm_bits) * table_scale;
    if (job >= b) {                                   // padding row of the query table: zeros
        const float z[4] = {0.f, 0.f, 0.f, 0.f};
        emit_folded_row(z, z, 2 * b + (job - b), -1, q_pad, qsplit, self_id, qscale, band, kappa, table_scale, table_norm, lane);
        return;
    }
    const long long i = job;
    const float *h = hr.row(i, kD), *t = tr.row(i, kD), *r = rr.row(i, kD);
    const float4 hv = __ldg(reinterpret_cast<const float4 *>(h) + lane), tv = __ldg(reinterpret_cast<const float4 *>(t) + lane),
                 rv = __ldg(reinterpret_cast<const float4 *>(r) + lane);
    const long long self_h = triples[i * 3 + 0], self_t = triples[i * 3 + 1];
    constexpr bool kHalves = MODEL != BLP_MODEL_DISTMULT;
    const bool first = !kHalves || lane < 16;          // this lane's four positions lie in [0, L)
    float4 hp = hv, tp = tv, rp = rv;                  // the same positions of the other half (halves models)
    if (kHalves) { hp = shfl_xor4(hv, 16); tp = shfl_xor4(tv, 16); rp = shfl_xor4(rv, 16); }
    const float ho[4] = {hv.x, hv.y, hv.z, hv.w}, to[4] = {tv.x, tv.y, tv.z, tv.w}, ro[4] = {rv.x, rv.y, rv.z, rv.w};
    const float hq[4] = {hp.x, hp.y, hp.z, hp.w}, tq[4] = {tp.x, tp.y, tp.z, tp.w}, rq[4] = {rp.x, rp.y, rp.z, rp.w};
#pragma unroll
    for (int role = 0; role < 2; ++role) {
        const bool head_pred = role == 0;
        float c[4], ca[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            // x[k] / x[L + k]: own / partner value in the first half, partner / own in the second
            const float hk = first ? ho[u] : hq[u], hLk = first ? hq[u] : ho[u];
            const float tk = first ? to[u] : tq[u], tLk = first ? tq[u] : to[u];
            const float rk = first ? ro[u] : rq[u], rLk = first ? rq[u] : ro[u];
            c[u] = fold_coeff_v<MODEL>(head_pred, first, hk, hLk, tk, tLk, rk, rLk);
            ca[u] = fold_abs_v<MODEL>(first, head_pred ? tk : hk, head_pred ? tLk : hLk, rk, rLk, c[u]);
        }
        emit_folded_row(c, ca, head_pred ? i : b + i, head_pred ? self_h : self_t, q_pad, qsplit, self_id, qscale, band, kappa,
                        table_scale, table_norm, lane);
    }
    if (true_score) {
        float s = true_score_warp128<MODEL>(h, t, r, s_terms[warp], lane);
        if (lane == 0) {
            if (!(hr.in_range(i) && tr.in_range(i) && rr.in_range(i))) s = __int_as_float(0x7fc00000);
            true_score[i] = s;
            true_score[tail_off + i] = s;
            gt[i] = 0; gt[tail_off + i] = 0;
            ge[i] = 0; ge[tail_off + i] = 0;
        }
    }
}

// unused tail [pos, end) of a warp's worklist block: skip entries
constexpr int kRefineBlock = 128;
__device__ __forceinline__ void refine_pad(const FastArgs &args, unsigned int pos, unsigned int end, int lane) {
    for (unsigned int p = pos + (unsigned int)lane; p < end; p += 32u)
        if ((long long)p < args.refine_cap) args.refine->entries[p] = make_int2(-1, 0);
}

// ---- the sweep --------------------------------------------------------------------------------------
template <bool PAIR>
__global__ void __launch_bounds__(kFThreads, 1) fast_sweep_kernel(const FastArgs args, const __grid_constant__ CUtensorMap tm_q,
                                                                  const __grid_constant__ CUtensorMap tm_e) {
    using Cfg = FastCfg<PAIR>;
    using FastSmem = FastSmemT<PAIR>;
    constexpr int kStages = Cfg::kStages;
    extern __shared__ unsigned char smem_raw[];
    FastSmem &sm = *reinterpret_cast<FastSmem *>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;     // 0 = leader: issues the MMAs, owns the `full` / d_empty barriers

    if (tid == 0) {
        mbar_init(&sm.a_full, 1);
        mbar_init(&sm.a_empty, 1);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&sm.b_full[s], 1);
            mbar_init(&sm.b_empty[s], 1);
        }
        for (int i = 0; i < kFBufs; ++i) {
            mbar_init(&sm.d_full[i], 1);
            mbar_init(&sm.d_empty[i], (PAIR ? 2 : 1) * kFEpiWarps);   // one arrival per epilogue warp (of both CTAs)
        }
        mbar_fence_init();
    }
    if (warp == 2) {
        if (PAIR) tmem_alloc2(&sm.tmem_base, kTmemCols);
        else tmem_alloc(&sm.tmem_base, kTmemCols);
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all();              // the peer's barriers are initialised before anything signals them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;

    // this CTA's share of the (query block, candidate tile) list, candidates fastest so the query block stays resident
    const long long total = args.m_tiles * args.n_tiles;
    const long long unit = PAIR ? blockIdx.x >> 1 : blockIdx.x, units = PAIR ? gridDim.x >> 1 : gridDim.x;   // both CTAs of a pair walk the same list
    const long long id_begin = total * unit / units, id_end = total * (unit + 1) / units;
    const long long q0_of_cta = (long long)rank * (kFH * kFM);       // this CTA's 256 queries inside the item's query block
    // 128-query halves of query block m that hold real queries (the last block may have one); a pair always runs both
    // (its MMAs span the two CTAs; padding rows are zeros and are never counted)
    auto halves_of = [&](long long m) -> int { return (PAIR || 2 * args.b - m * (kFH * kFM) > kFM) ? 2 : 1; };

    if (warp == 0) {
        if (lane == 0) {
            // ===================== TMA producer =====================
            long long cur_m = -1;
            uint32_t a_use = 0, kbit = 0;
            // (query block m, candidate tile n) walk incrementally: a 64-bit division per item and role is ~100 instructions
            long long m = id_begin / args.n_tiles, n = id_begin % args.n_tiles - 1;
            for (long long id = id_begin; id < id_end; ++id) {
                if (++n == args.n_tiles) { n = 0; ++m; }
                if (m != cur_m) {
                    mbar_wait(&sm.a_empty, (a_use & 1u) ^ 1u);
                    // the leader's barrier collects the bytes of both CTAs' query blocks
                    if (rank == 0) mbar_arrive_expect_tx(&sm.a_full, (PAIR ? 2 : 1) * kFH * 2 * 2 * kFBox * 2);
#pragma unroll
                    for (int half = 0; half < kFH; ++half)
#pragma unroll
                        for (int hl = 0; hl < 2; ++hl)
#pragma unroll
                            for (int kb = 0; kb < 2; ++kb) {
                                const int qrow = (int)(hl * args.q_pad + m * Cfg::kQ + q0_of_cta + half * kFM);
                                if (PAIR) tma_tensor2d_g2s_pair(&sm.a[half][hl][kb][0], &tm_q, kb * kFKB, qrow, &sm.a_full);
                                else tma_tensor2d_g2s(&sm.a[half][hl][kb][0], &tm_q, kb * kFKB, qrow, &sm.a_full);
                            }
                    ++a_use;
                    cur_m = m;
                }
                for (int kb = 0; kb < 2; ++kb, ++kbit) {
                    const int stage = kbit % kStages;
                    const uint32_t use = kbit / kStages;
                    mbar_wait(&sm.b_empty[stage], (use & 1u) ^ 1u);
                    if (!PAIR && (args.debug & 2) && kbit >= kStages) { mbar_arrive(&sm.b_full[stage]); continue; }
                    if (rank == 0) mbar_arrive_expect_tx(&sm.b_full[stage], (PAIR ? 2 : 1) * 2 * Cfg::kBoxB * 2);
                    const int crow = (int)(n * kFN + rank * Cfg::kRowsB);     // a pair splits the 128 candidate rows
                    if (PAIR) {
                        tma_tensor2d_g2s_pair(&sm.b[stage][0][0], &tm_e, kb * kFKB, crow, &sm.b_full[stage]);
                        tma_tensor2d_g2s_pair(&sm.b[stage][1][0], &tm_e, kb * kFKB, (int)args.n_pad + crow, &sm.b_full[stage]);
                    } else {
                        tma_tensor2d_g2s(&sm.b[stage][0][0], &tm_e, kb * kFKB, crow, &sm.b_full[stage]);
                        tma_tensor2d_g2s(&sm.b[stage][1][0], &tm_e, kb * kFKB, (int)args.n_pad + crow, &sm.b_full[stage]);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1 && rank == 0) {
        // ===================== MMA issuer (the leader CTA of a pair) =====================
        // The WHOLE warp runs this loop in convergent code and one elected lane issues the tcgen05 instructions: every
        // operand (descriptors, TMEM addresses) is then provably warp-uniform and lives in uniform registers.  With the
        // loop under `if (lane == 0)` ptxas had to move each operand into a uniform register through an ELECT /
        // R2UR.BROADCAST / BRA.U.ANY waterfall -- 13 dependent instructions in front of every UTCHMMA, 118 clocks per
        // MMA issued against the 64-clock floor of an M128 x N128 x K16 MMA (profiles/r02_fast_kernel_experiments.txt).
        constexpr uint32_t idesc = umma_idesc_f16(PAIR ? 2 * kFM : kFM, kFN);
        long long cur_m = -1;
        uint32_t a_use = 0, kbit = 0, it = 0;
        int halves = kFH;
        long long m = id_begin / args.n_tiles, n = id_begin % args.n_tiles - 1;
        for (long long id = id_begin; id < id_end; ++id, ++it) {
            if (++n == args.n_tiles) { n = 0; ++m; }
            if (m != cur_m) {
                mbar_wait(&sm.a_full, a_use & 1u);
                ++a_use;
                cur_m = m;
                halves = halves_of(m);
            }
            const uint32_t buf = it % kFBufs, duse = it / kFBufs;
            mbar_wait(&sm.d_empty[buf], (duse & 1u) ^ 1u);          // epilogue has drained these accumulators
            tc_fence_after();
            for (int kb = 0; kb < 2; ++kb, ++kbit) {
                const int stage = kbit % kStages;
                const uint32_t use = kbit / kStages;
                mbar_wait(&sm.b_full[stage], use & 1u);
                tc_fence_after();
                const uint64_t b_hi = umma_smem_desc(&sm.b[stage][0][0]), b_lo = umma_smem_desc(&sm.b[stage][1][0]);
                for (int half = 0; half < ((args.debug & 4) ? 0 : halves); ++half) {   // the candidate k-block serves both query halves
                    const uint32_t d_tmem = tmem + buf * (kFH * kFN) + half * kFN;
                    const uint64_t a_hi = umma_smem_desc(&sm.a[half][0][kb][0]), a_lo = umma_smem_desc(&sm.a[half][1][kb][0]);
                    if (elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < kFKB / 16; ++ks) {        // 16 halves = 32 bytes = 2 descriptor units per step
                            const uint64_t o = (uint64_t)(ks * 2);
                            if (PAIR) {
                                umma_f16_2(d_tmem, a_lo + o, b_hi + o, idesc, (kb | ks) != 0);
                                umma_f16_2(d_tmem, a_hi + o, b_lo + o, idesc, 1u);
                                umma_f16_2(d_tmem, a_hi + o, b_hi + o, idesc, 1u);
                            } else {
                                umma_f16(d_tmem, a_lo + o, b_hi + o, idesc, (kb | ks) != 0);
                                umma_f16(d_tmem, a_hi + o, b_lo + o, idesc, 1u);
                                umma_f16(d_tmem, a_hi + o, b_hi + o, idesc, 1u);
                            }
                        }
                    }
                    __syncwarp();
                }
                if (elect_one()) {                                  // frees the stage (in both CTAs) once these MMAs have read it
                    if (PAIR) umma_commit2(&sm.b_empty[stage]);
                    else umma_commit(&sm.b_empty[stage]);
                }
                __syncwarp();
            }
            if (elect_one()) {
                const bool last_of_block = id + 1 == id_end || n + 1 == args.n_tiles;
                if (PAIR) {
                    umma_commit2(&sm.d_full[buf]);
                    if (last_of_block) umma_commit2(&sm.a_empty);
                } else {
                    umma_commit(&sm.d_full[buf]);
                    if (last_of_block) umma_commit(&sm.a_empty);
                }
            }
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ===================== epilogue: one query row of each half per thread =====================
        // TMEM lane = query row; a warp can only read the lane quarter (warp % 4), so the two warps of a quarter
        // split the 32-column chunks between them (chunk parity = part)
        const int ew = warp & 3, part = (warp - 4) >> 2, row = ew * 32 + lane;
        long long cur_m = -1;
        long long slot[kFH], self_local[kFH];
        float th[kFH], inv[kFH], bd[kFH];                                // scaled true score, 1 / scale, refine band
        bool valid_q[kFH];
        int cgt[kFH], cge[kFH];
        int halves = kFH;
#pragma unroll
        for (int h = 0; h < kFH; ++h) { slot[h] = -1; self_local[h] = -1; th[h] = 0.f; inv[h] = 1.f; bd[h] = 0.f; valid_q[h] = false; cgt[h] = cge[h] = 0; }
        uint32_t it = 0;
        unsigned int wl_pos = 0u, wl_end = 0u;                           // this warp's block of worklist slots (refine mode)
        auto flush = [&]() {
#pragma unroll
            for (int h = 0; h < kFH; ++h) {
                if (valid_q[h] && (cgt[h] | cge[h])) {
                    atomicAdd(&args.gt[slot[h]], cgt[h]);
                    atomicAdd(&args.ge[slot[h]], cge[h]);
                }
                cgt[h] = cge[h] = 0;
            }
        };
        long long m = id_begin / args.n_tiles, n = id_begin % args.n_tiles - 1;
        const bool want_scores = args.scores_out != nullptr, refine = args.refine != nullptr;
        for (long long id = id_begin; id < id_end; ++id, ++it) {
            if (++n == args.n_tiles) { n = 0; ++m; }
            if (m != cur_m) {
                flush();
                halves = halves_of(m);
#pragma unroll
                for (int h = 0; h < kFH; ++h) {
                    const long long q = m * Cfg::kQ + q0_of_cta + h * kFM + row;
                    valid_q[h] = q < 2 * args.b;
                    if (valid_q[h]) {
                        slot[h] = q < args.b ? q : args.tail_off + (q - args.b);
                        const float st = args.true_score[slot[h]];
                        const float qs = args.qscale[q];
                        th[h] = fmul(st, qs);                            // power-of-two scale: exact
                        inv[h] = __frcp_rn(qs);
                        bd[h] = args.band ? args.band[q] : 0.f;
                        self_local[h] = args.self_id[q] - args.ent_offset;
                        valid_q[h] = st == st;                           // NaN = flagged triple (bad index)
                    }
                }
                cur_m = m;
            }
            const uint32_t buf = it % kFBufs, duse = it / kFBufs;
            mbar_wait(&sm.d_full[buf], duse & 1u);
            tc_fence_after();
            const long long tile_base = n * kFN;
            const int nvalid = (int)min((long long)kFN, args.n_local - tile_base);
#pragma unroll
            for (int h = 0; h < kFH; ++h) {
                if (h >= halves || (args.debug & 1)) continue;           // uniform over the CTA
                // column of the true entity if it lives in this tile (else out of range)
                const long long self_off = self_local[h] - tile_base;
                const int self_col = (!refine && self_off >= 0 && self_off < nvalid) ? (int)self_off : -1;
                const uint32_t taddr = tmem + ((uint32_t)(ew * 32) << 16) + buf * (kFH * kFN) + h * kFN;
                const float st = th[h];
                const long long q = m * Cfg::kQ + q0_of_cta + h * kFM + row;
#pragma unroll 1
                for (int ch = part; ch < kFN / 32; ch += kFEpiWarps / 4) {
                    uint32_t v[32];
                    __syncwarp();                                        // the TMEM load is warp-collective
                    tmem_ld32(taddr + ch * 32, v);
                    tmem_ld_wait();
                    if (want_scores && q < 2 * args.b) {
                        float *orow = args.scores_out + q * args.ld_scores + tile_base + ch * 32;
#pragma unroll
                        for (int c = 0; c < 32; ++c)
                            if (ch * 32 + c < nvalid) orow[c] = fmul(__uint_as_float(v[c]), inv[h]);
                    }
                    // plain fast mode: bd == 0, both thresholds are the scaled true score.  Refine mode: `gs` counts the
                    // candidates that beat the true score for certain (s > st + band), `es` those that may tie or beat
                    // it (s >= st - band); the difference is the band, which goes to the worklist
                    const float st_hi = st + bd[h], st_lo = st - bd[h];
                    int gs = 0, es = 0;
                    if (valid_q[h]) {
                        int g[4] = {0, 0, 0, 0}, e[4] = {0, 0, 0, 0};    // independent chains
                        if (nvalid == kFN) {
#pragma unroll
                            for (int c = 0; c < 32; ++c) {
                                const float s = __uint_as_float(v[c]);
                                g[c & 3] += s > st_hi;
                                e[c & 3] += s >= st_lo;
                            }
                        } else {
#pragma unroll
                            for (int c = 0; c < 32; ++c) {
                                const float s = __uint_as_float(v[c]);
                                const bool ok = ch * 32 + c < nvalid;
                                g[c & 3] += ok && s > st_hi;
                                e[c & 3] += ok && s >= st_lo;
                            }
                        }
                        gs = (g[0] + g[1]) + (g[2] + g[3]);
                        es = (e[0] + e[1]) + (e[2] + e[3]);
                        if ((self_col >> 5) == ch) {
                            // the true entity itself: it ties with s_true by definition (utils.py:104-105), whatever
                            // the fast arithmetic produced for it
                            const float s_self = __uint_as_float(select32(v, (int)(self_col & 31)));
                            gs -= s_self > st;
                            es += 1 - (s_self >= st);
                        }
                    }
                    if (refine) {                                        // warp-uniform
                        // The band of this 32-column chunk goes to the worklist.  Slots are handed out warp-wide: the
                        // warp owns a block of kRefineBlock entries (ONE global atomic per block -- a returning atomic
                        // per entry on a single address serialises the whole chip: 0.33 -> 1.1 ms per 16,384 triples),
                        // a ballot ranks the lanes that append at the same column; what is left of a block is padded
                        // with skip entries (q = -1).
                        const bool mine = es != gs;
                        if (__any_sync(0xffffffffu, mine)) {             // ~1 chunk visit in 5 at 3 band candidates per query
                            // per-lane bit mask of the band columns (branch-free, small), then a ROLLED loop over the
                            // columns that have a band candidate in any lane -- usually one.  (Unrolling the append 32
                            // times made this rare path ~50 KB of code: every visit thrashed the instruction cache,
                            // 3 -> 17 us per work item.)
                            unsigned int bm = 0u;
#pragma unroll
                            for (int c = 0; c < 32; ++c) {
                                const float s = __uint_as_float(v[c]);
                                bm |= (s >= st_lo && !(s > st_hi)) ? (1u << c) : 0u;
                            }
                            if (!mine) bm = 0u;
                            const int left = nvalid - ch * 32;           // valid columns of this chunk
                            if (left < 32) bm &= left <= 0 ? 0u : ((1u << left) - 1u);
                            unsigned int cols = __reduce_or_sync(0xffffffffu, bm);
#pragma unroll 1
                            while (cols) {
                                const int c = __ffs((int)cols) - 1;
                                cols &= cols - 1u;
                                const bool in = (bm >> c) & 1u;
                                const unsigned int mask = __ballot_sync(0xffffffffu, in);
                                const unsigned int k = (unsigned int)__popc(mask);
                                if (wl_pos + k > wl_end) {
                                    refine_pad(args, wl_pos, wl_end, lane);
                                    unsigned int base = 0u;
                                    if (lane == 0) {
                                        // far past any capacity (< 2^31): stop counting so the 32-bit counter cannot wrap
                                        if (*reinterpret_cast<volatile unsigned int *>(&args.refine->count) >= 0xC0000000u) {
                                            args.refine->overflow = 1u;
                                            base = 0xC0000000u;
                                        } else {
                                            base = atomicAdd(&args.refine->count, (unsigned int)kRefineBlock);
                                        }
                                    }
                                    wl_pos = __shfl_sync(0xffffffffu, base, 0);
                                    wl_end = wl_pos + kRefineBlock;
                                }
                                if (in) {
                                    const unsigned int at = wl_pos + (unsigned int)__popc(mask & ((1u << lane) - 1u));
                                    if ((long long)at < args.refine_cap)
                                        args.refine->entries[at] = make_int2((int)q, (int)(tile_base + ch * 32 + c));
                                }
                                wl_pos += k;
                            }
                        }
                        es = gs;                                         // the refine kernel adds the band's exact verdicts
                    }
                    cgt[h] += gs;
                    cge[h] += es;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {                                             // the leader's MMA warp waits for both CTAs' epilogues
                if (PAIR) mbar_arrive_leader(&sm.d_empty[buf]);
                else mbar_arrive(&sm.d_empty[buf]);
            }
        }
        flush();
        if (args.refine) refine_pad(args, wl_pos, wl_end, lane);
    }

    tc_fence_before();
    if (PAIR) cluster_sync_all();              // the peer may still be read by the pair's MMAs / signalled by this CTA
    else __syncthreads();
    if (warp == 2) {
        if (PAIR) tmem_dealloc2(tmem, kTmemCols);
        else tmem_dealloc(tmem, kTmemCols);
    }
}

// ---- refine pass: exact verdicts for the band -----------------------------------------------------------
// Every worklist entry (query q, candidate row) is re-scored with the reference's own fp32 operations and summation
// order (true_scores_warp128x4: the same code that produces the true-triple scores, so the true entity ties itself
// bit for bit) and compared with the exact true score; four entries per warp, their summation chains side by side.
// With the certain counts of the sweep kernel this makes gt / ge equal to the exact mode's, as long as the band
// really bounds |fast - reference| (tests/test_gpu_fast.py compares the two modes bit for bit).
constexpr int kRefineWarps = 8;
template <int MODEL>
__global__ void __launch_bounds__(kRefineWarps * 32) refine_kernel(RefineList *__restrict__ wl, long long cap,
                                                                   const float *__restrict__ ent, const RowRef hr,
                                                                   const RowRef tr, const RowRef rr, long long b,
                                                                   long long tail_off, const float *__restrict__ true_score,
                                                                   int *__restrict__ gt, int *__restrict__ ge) {
    __shared__ __align__(16) float tm[kRefineWarps][4 * kD];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned int count = *reinterpret_cast<volatile unsigned int *>(&wl->count);
    const long long n = (long long)count < cap ? (long long)count : cap;
    if (blockIdx.x == 0 && threadIdx.x == 0 && (long long)count > cap) wl->overflow = 1u;   // entries were dropped: results invalid
    const long long stride = (long long)gridDim.x * kRefineWarps * 4;
    for (long long base = ((long long)blockIdx.x * kRefineWarps + warp) * 4; base < n; base += stride) {
        const float *h[4], *t[4], *r[4];
        long long slot[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            h[u] = t[u] = r[u] = ent;                                      // idle job (past the end / skip entry): any valid row
            slot[u] = -1;
            const int2 e = base + u < n ? wl->entries[base + u] : make_int2(-1, 0);
            if (e.x >= 0) {
                const long long q = e.x;
                const bool head_pred = q < b;
                const long long i = head_pred ? q : q - b;
                const float *c = ent + (long long)e.y * kD;
                slot[u] = head_pred ? i : tail_off + i;
                h[u] = head_pred ? c : hr.row(i, kD);
                t[u] = head_pred ? tr.row(i, kD) : c;
                r[u] = rr.row(i, kD);
            }
        }
        if (slot[0] < 0 && slot[1] < 0 && slot[2] < 0 && slot[3] < 0) continue;     // warp-uniform: a padded block tail
        const float s = true_scores_warp128x4<MODEL>(h, t, r, 4, tm[warp], lane);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (slot[u] >= 0 && lane == true_job_lane<MODEL>(u)) {
                const float st = true_score[slot[u]];
                if (s > st) atomicAdd(gt + slot[u], 1);
                if (s >= st) atomicAdd(ge + slot[u], 1);
            }
        }
    }
}

// ---- host side -----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn fast_encode_fn() {
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
// [rows][128 halves] viewed as boxes of [box_rows rows][64 halves], 128-byte swizzle
static bool make_rows_tmap(CUtensorMap *tm, const void *base, long long rows, int box_rows) {
    EncodeTiledFn enc = fast_encode_fn();
    if (!enc) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)kD, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)kD * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kFKB, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }

// table: hi + lo halves of n_pad rows, then the header
long long fast_table_ws_bytes(long long n_local) { return 2 * round_up(n_local > 0 ? n_local : 1, kFN) * kD * 2 + 256; }
// queries: hi + lo halves of q_pad rows, self ids, scales
long long fast_query_ws_bytes(long long t) {
    const long long q_pad = round_up(2 * (t > 0 ? t : 1), 2 * kFH * kFM);      // a CTA pair's query block
    return 2 * q_pad * kD * 2 + q_pad * 8 + q_pad * 4 + q_pad * 4;     // + the refine band per query row
}
// worklist of the refine pass: header + capacity (query, candidate) pairs
long long fast_refine_ws_bytes(long long capacity) { return 16 + 8 * (capacity > 0 ? capacity : 1); }

int fast_prepare_table(const float *ent, long long n_local, void *table_ws, cudaStream_t st) {
    const long long n_pad = round_up(n_local > 0 ? n_local : 1, kFN);
    __half *out = reinterpret_cast<__half *>(table_ws);
    FastTableHeader *hdr = reinterpret_cast<FastTableHeader *>(out + 2 * n_pad * kD);
    BLP_CUDA(cudaMemsetAsync(hdr, 0, sizeof(FastTableHeader), st));
    const long long total4 = n_local * (kD / 4);
    if (total4 > 0) {
        long long blocks = (total4 + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        table_maxabs_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float4 *>(ent), total4, &hdr->max_bits);
        count_launch();
        BLP_CUDA(cudaGetLastError());
    }
    if (n_local > 0) {
        long long blocks = (n_local * 32 + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        table_maxnorm_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float4 *>(ent), n_local, &hdr->max_norm_bits);
        count_launch();
        BLP_CUDA(cudaGetLastError());
    }
    const long long total = n_pad * (kD / 4);
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    split_table_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float4 *>(ent), n_local, n_pad, out, hdr);
    count_launch();
    return check_cuda(cudaGetLastError(), "split_table_kernel launch");
}

static int num_sms_fast() {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
}

// Folds the queries, then counts gt / ge over this shard with the tcgen05 kernel.  With compute_true the fold kernel
// also writes the exact true scores and zeroes the counters; otherwise gt / ge / true_score must already hold zeros /
// the exact true scores (a true-score kernel has run on the same stream).
int launch_fast_sweep(int model, long long n_local, long long ent_offset, const RowRef &h, const RowRef &t, const RowRef &r,
                      const long long *triples, long long b, long long tail_off, float *true_score, int *gt, int *ge,
                      const void *table_ws, void *query_ws, float *scores_out, long long ld_scores, bool compute_true,
                      const float *ent, void *refine_ws, long long refine_cap, cudaStream_t st) {
    if (model != BLP_MODEL_DISTMULT && model != BLP_MODEL_COMPLEX && model != BLP_MODEL_SIMPLE) {
        set_error("fast (tensor-core) mode covers distmult / complex / simple; transe is an L1 distance, not a contraction");
        return BLP_EINVAL;
    }
    FastArgs a{};
    a.n_local = n_local; a.ent_offset = ent_offset; a.n_pad = round_up(n_local, kFN);
    a.b = b; a.tail_off = tail_off; a.q_pad = round_up(2 * b, 2 * kFH * kFM);
    // CTA pairs (cta_group::2) from 1,024 queries on: below that the 512-query blocks of a pair leave SMs idle
    static const int pair_env = []() {
        const char *e = getenv("BLP_FAST_PAIR");          // experiments: 0 / 1 forces the choice
        return e ? atoi(e) : -1;
    }();
    const long long sms = num_sms_fast();
    a.pair = (pair_env >= 0 ? pair_env != 0 : 2 * b >= 1024) && sms >= 2 ? 1 : 0;
    const long long qblock = (a.pair ? 2 : 1) * kFH * kFM;
    a.m_tiles = a.q_pad / qblock; a.n_tiles = a.n_pad / kFN;
    __half *qsplit = reinterpret_cast<__half *>(query_ws);
    long long *self_id = reinterpret_cast<long long *>(qsplit + 2 * a.q_pad * kD);
    float *qscale = reinterpret_cast<float *>(self_id + a.q_pad);
    float *band = qscale + a.q_pad;
    // refine band: kappa bounds |fast - reference| relative to ||fold_abs||_2 * max ||e||_2
    static const float kappa = []() {
        const char *e = getenv("BLP_FAST_KAPPA");          // experiments only
        const float v = e ? (float)atof(e) : 0.0f;
        return v > 0.0f ? v : 2.0e-5f;
    }();
    if (refine_ws) {
        a.band = band;
        a.refine = reinterpret_cast<RefineList *>(refine_ws);
        a.refine_cap = refine_cap;
        BLP_CUDA(cudaMemsetAsync(&a.refine->count, 0, sizeof(unsigned int), st));
    }
    const __half *table = reinterpret_cast<const __half *>(table_ws);
    const FastTableHeader *hdr = reinterpret_cast<const FastTableHeader *>(table + 2 * a.n_pad * kD);
    a.true_score = true_score; a.self_id = self_id; a.qscale = qscale; a.gt = gt; a.ge = ge;
    a.scores_out = scores_out; a.ld_scores = ld_scores;
    {
        const char *e = getenv("BLP_FAST_DEBUG");
        a.debug = e ? atoi(e) : 0;
    }

    float *ts_out = compute_true ? true_score : nullptr;
    // one warp per triple when the rows allow 16-byte loads (every d = 128 table / dense row block that is 16-byte aligned)
    const bool warp_fold = ((reinterpret_cast<uintptr_t>(h.base) | reinterpret_cast<uintptr_t>(t.base) | reinterpret_cast<uintptr_t>(r.base)) & 15u) == 0;
    const unsigned fold_grid = (unsigned)((b + (a.q_pad - 2 * b) + kFoldWarps - 1) / kFoldWarps);
#define BLP_FOLD(M)                                                                                                            \
    do {                                                                                                                       \
        if (warp_fold)                                                                                                         \
            fold_triples_kernel<M><<<fold_grid, kFoldWarps * 32, 0, st>>>(h, t, r, triples, b, a.q_pad, qsplit, self_id, qscale, hdr, \
                                                                          tail_off, ts_out, gt, ge, band, kappa);            \
        else                                                                                                                   \
            fold_queries_kernel<M><<<(unsigned)a.q_pad, kD, 0, st>>>(h, t, r, triples, b, a.q_pad, qsplit, self_id, qscale, hdr,  \
                                                                     tail_off, ts_out, gt, ge, band, kappa);                 \
    } while (0)
    switch (model) {
    case BLP_MODEL_DISTMULT: BLP_FOLD(BLP_MODEL_DISTMULT); break;
    case BLP_MODEL_COMPLEX: BLP_FOLD(BLP_MODEL_COMPLEX); break;
    default: BLP_FOLD(BLP_MODEL_SIMPLE); break;
    }
#undef BLP_FOLD
    count_launch();
    BLP_CUDA(cudaGetLastError());

    CUtensorMap tm_q, tm_e;
    memset(&tm_q, 0, sizeof(tm_q));
    memset(&tm_e, 0, sizeof(tm_e));
    if (!make_rows_tmap(&tm_q, qsplit, 2 * a.q_pad, kFM) ||
        !make_rows_tmap(&tm_e, table_ws, 2 * a.n_pad, a.pair ? kFN / 2 : kFN)) {
        set_error("cuTensorMapEncodeTiled failed for the fast-mode operand tables");
        return BLP_ECUDA;
    }
    const long long items = a.m_tiles * a.n_tiles;
    if (items == 0) return BLP_OK;
    if (a.pair) {
        const size_t smem = sizeof(FastSmemT<true>) + 1024;
        BLP_CUDA(cudaFuncSetAttribute(fast_sweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const long long pairs = sms / 2;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(2 * (items < pairs ? items : pairs)));
        cfg.blockDim = dim3(kFThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        prof_begin(1, st);
        BLP_CUDA(cudaLaunchKernelEx(&cfg, fast_sweep_kernel<true>, a, tm_q, tm_e));
        prof_end(1, st);
    } else {
        const size_t smem = sizeof(FastSmemT<false>) + 1024;
        BLP_CUDA(cudaFuncSetAttribute(fast_sweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const unsigned grid = (unsigned)(items < sms ? items : sms);
        prof_begin(1, st);
        fast_sweep_kernel<false><<<grid, kFThreads, smem, st>>>(a, tm_q, tm_e);
        prof_end(1, st);
    }
    count_launch();
    BLP_CUDA(cudaGetLastError());
    if (refine_ws) {
        const unsigned rgrid = (unsigned)(sms * 4);
        switch (model) {
        case BLP_MODEL_DISTMULT: refine_kernel<BLP_MODEL_DISTMULT><<<rgrid, kRefineWarps * 32, 0, st>>>(a.refine, refine_cap, ent, h, t, r, b, tail_off, true_score, gt, ge); break;
        case BLP_MODEL_COMPLEX: refine_kernel<BLP_MODEL_COMPLEX><<<rgrid, kRefineWarps * 32, 0, st>>>(a.refine, refine_cap, ent, h, t, r, b, tail_off, true_score, gt, ge); break;
        default: refine_kernel<BLP_MODEL_SIMPLE><<<rgrid, kRefineWarps * 32, 0, st>>>(a.refine, refine_cap, ent, h, t, r, b, tail_off, true_score, gt, ge); break;
        }
        count_launch();
        BLP_CUDA(cudaGetLastError());
    }
    return BLP_OK;
}

}  // namespace blp
