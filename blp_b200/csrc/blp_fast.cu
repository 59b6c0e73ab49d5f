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
// Arithmetic: 3xTF32.  Every fp32 operand x is split as x = hi + lo with hi = tf32(x), lo = tf32(x - hi)
// (cvt.rna), and hi*hi + lo*hi + hi*lo is accumulated in fp32 in TMEM: ~2^-21 relative per product.
//
// Kernel (sm_100a, one persistent CTA per SM, 256 threads):
//   warp 0 lane 0   TMA producer: query tile (hi | lo, 128 KB, resident per M tile) and a 3-stage ring of
//                   candidate K-blocks (hi | lo boxes of 128 rows x 32 floats, 128-byte swizzle)
//   warp 1 lane 0   tcgen05.mma issuer (kind::tf32, M = 128 queries, N = 128 candidates, K = 8 per
//                   instruction, 48 instructions per tile), accumulators double-buffered in TMEM
//   warp 2          TMEM allocation (256 columns)
//   warps 4-7       epilogue: tcgen05.ld 32 columns at a time -- TMEM lane = query row, so each thread owns
//                   one query -- compare against the true score, count, one atomicAdd pair per query and
//                   M-tile run.  No score ever leaves the SM.
#include <cuda.h>
#include <string.h>

#include "blp_sweep.h"

namespace blp {

constexpr int kFM = 128, kFN = 128;
constexpr int kFStages = 3;
constexpr int kFBox = 128 * 32;            // floats in one TMA box (128 rows x 32 floats = 16 KB)
constexpr int kFThreads = 256;
constexpr int kTmemCols = 256;             // two 128-column accumulators

struct __align__(1024) FastSmem {
    float a[2][4][kFBox];                  // [hi | lo][k-block] query tile
    float b[kFStages][2][kFBox];           // [stage][hi | lo] candidate k-block
    uint64_t a_full, a_empty;
    uint64_t b_full[kFStages], b_empty[kFStages];
    uint64_t d_full[2], d_empty[2];
    uint32_t tmem_base;
};

struct FastArgs {
    long long n_local, ent_offset, n_pad;  // candidates in this shard, global id of row 0, rows per half of the split table
    long long b, tail_off, q_pad;          // triples, output slot offset of tail queries, rows per half of the split queries
    long long m_tiles, n_tiles;
    const float *true_score;               // indexed by output slot
    const long long *self_id;              // [q_pad] global candidate id of the query's true entity (-1 = padding)
    int *gt, *ge;
    float *scores_out;                     // optional (2b, ld_scores) matrix of the fast scores (verification aid)
    long long ld_scores;
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
// D[tmem] (+)= A[smem desc] . B[smem desc]^T, tf32 inputs, fp32 accumulation
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
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

// K-major operand tile of [rows][32 floats] written by TMA with the 128-byte swizzle: 8-row groups are
// 1024 bytes apart (SBO), the leading-dimension offset is unused for swizzled K-major layouts.
__device__ __forceinline__ uint64_t umma_smem_desc(const void *p) {
    uint64_t d = (uint64_t)((smem_u32(p) & 0x3FFFFu) >> 4);   // start address, 16-byte units
    d |= (uint64_t)1 << 16;                                    // leading byte offset (ignored)
    d |= (uint64_t)(1024 >> 4) << 32;                          // stride byte offset
    d |= (uint64_t)1 << 46;                                    // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                                    // SWIZZLE_128B
    return d;
}
// kind::tf32, fp32 accumulator, both operands K-major, M x N
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
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
// Split table: rows [0, n_pad) = tf32(e), rows [n_pad, 2 n_pad) = tf32(e - tf32(e)); rows >= n_local are zero.
__global__ void split_table_kernel(const float4 *__restrict__ ent, long long n_local, long long n_pad, float4 *__restrict__ out) {
    const long long total = n_pad * (kD / 4);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / (kD / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < n_local) v = __ldg(ent + i);
        float4 hi, lo;
        hi.x = tf32_rna(v.x); hi.y = tf32_rna(v.y); hi.z = tf32_rna(v.z); hi.w = tf32_rna(v.w);
        lo.x = tf32_rna(fsub(v.x, hi.x)); lo.y = tf32_rna(fsub(v.y, hi.y));
        lo.z = tf32_rna(fsub(v.z, hi.z)); lo.w = tf32_rna(fsub(v.w, hi.w));
        out[i] = hi;
        out[total + i] = lo;
    }
}

// Coefficient of candidate element j for query (triple i, role): the score with the candidate factored out.
template <int MODEL>
__device__ __forceinline__ float fold_coeff(bool head_pred, const float *__restrict__ h, const float *__restrict__ t,
                                            const float *__restrict__ r, int j) {
    constexpr int L = kD / 2;
    if (MODEL == BLP_MODEL_DISTMULT) return head_pred ? fmul(r[j], t[j]) : fmul(h[j], r[j]);   // models.py:227
    const int k = j & (L - 1);
    const bool first = j < L;
    if (MODEL == BLP_MODEL_COMPLEX) {   // models.py:230-239; r = (rr | ri), h = (hr | hi), t = (tr | ti)
        const float rr = r[k], ri = r[L + k];
        if (head_pred) {                // candidate = h:  hr (rr tr + ri ti) + hi (rr ti - ri tr)
            const float tr = t[k], ti = t[L + k];
            return first ? fadd(fmul(rr, tr), fmul(ri, ti)) : fsub(fmul(rr, ti), fmul(ri, tr));
        }
        const float hr = h[k], hi = h[L + k];   // candidate = t:  tr (rr hr - ri hi) + ti (rr hi + ri hr)
        return first ? fsub(fmul(rr, hr), fmul(ri, hi)) : fadd(fmul(rr, hi), fmul(ri, hr));
    }
    // SIMPLE, models.py:242-248: 1/2 (hh ra tt + th rb ht); h = (hh | ht), t = (th | tt), r = (ra | rb)
    if (head_pred) return first ? fmul(0.5f, fmul(r[k], t[L + k])) : fmul(0.5f, fmul(t[k], r[L + k]));
    return first ? fmul(0.5f, fmul(r[L + k], h[L + k])) : fmul(0.5f, fmul(h[k], r[k]));
}

// One CTA of 128 threads per query row: rows [0, b) predict heads, [b, 2b) predict tails, the rest is padding.
template <int MODEL>
__global__ void __launch_bounds__(kD) fold_queries_kernel(const RowRef hr, const RowRef tr, const RowRef rr,
                                                          const long long *__restrict__ triples, long long b,
                                                          long long q_pad, float *__restrict__ qsplit,
                                                          long long *__restrict__ self_id) {
    const long long q = blockIdx.x;
    const int j = threadIdx.x;
    float c = 0.0f;
    long long self = -1;
    if (q < 2 * b) {
        const bool head_pred = q < b;
        const long long i = head_pred ? q : q - b;
        c = fold_coeff<MODEL>(head_pred, hr.row(i, kD), tr.row(i, kD), rr.row(i, kD), j);
        self = triples[i * 3 + (head_pred ? 0 : 1)];
    }
    const float hi = tf32_rna(c);
    qsplit[q * kD + j] = hi;
    qsplit[(q_pad + q) * kD + j] = tf32_rna(fsub(c, hi));
    if (j == 0) self_id[q] = self;
}

// ---- the sweep --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFThreads, 1) fast_sweep_kernel(const FastArgs args, const __grid_constant__ CUtensorMap tm_q,
                                                                  const __grid_constant__ CUtensorMap tm_e) {
    extern __shared__ unsigned char smem_raw[];
    FastSmem &sm = *reinterpret_cast<FastSmem *>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        mbar_init(&sm.a_full, 1);
        mbar_init(&sm.a_empty, 1);
        for (int s = 0; s < kFStages; ++s) {
            mbar_init(&sm.b_full[s], 1);
            mbar_init(&sm.b_empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sm.d_full[i], 1);
            mbar_init(&sm.d_empty[i], 4);        // one arrival per epilogue warp
        }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(&sm.tmem_base, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;

    // this CTA's share of the (M tile, N tile) list, N fastest so the query tile stays resident
    const long long total = args.m_tiles * args.n_tiles;
    const long long id_begin = total * blockIdx.x / gridDim.x, id_end = total * (blockIdx.x + 1) / gridDim.x;

    if (warp == 0) {
        if (lane == 0) {
            // ===================== TMA producer =====================
            long long cur_m = -1;
            uint32_t a_use = 0, kbit = 0;
            for (long long id = id_begin; id < id_end; ++id) {
                const long long m = id / args.n_tiles, n = id % args.n_tiles;
                if (m != cur_m) {
                    mbar_wait(&sm.a_empty, (a_use & 1u) ^ 1u);
                    mbar_arrive_expect_tx(&sm.a_full, 2 * 4 * kFBox * 4);
#pragma unroll
                    for (int hl = 0; hl < 2; ++hl)
#pragma unroll
                        for (int kb = 0; kb < 4; ++kb)
                            tma_tensor2d_g2s(&sm.a[hl][kb][0], &tm_q, kb * 32, (int)(hl * args.q_pad + m * kFM), &sm.a_full);
                    ++a_use;
                    cur_m = m;
                }
                for (int kb = 0; kb < 4; ++kb, ++kbit) {
                    const int stage = kbit % kFStages;
                    const uint32_t use = kbit / kFStages;
                    mbar_wait(&sm.b_empty[stage], (use & 1u) ^ 1u);
                    mbar_arrive_expect_tx(&sm.b_full[stage], 2 * kFBox * 4);
                    tma_tensor2d_g2s(&sm.b[stage][0][0], &tm_e, kb * 32, (int)(n * kFN), &sm.b_full[stage]);
                    tma_tensor2d_g2s(&sm.b[stage][1][0], &tm_e, kb * 32, (int)(args.n_pad + n * kFN), &sm.b_full[stage]);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            // ===================== MMA issuer =====================
            constexpr uint32_t idesc = umma_idesc_tf32(kFM, kFN);
            long long cur_m = -1;
            uint32_t a_use = 0, kbit = 0, it = 0;
            for (long long id = id_begin; id < id_end; ++id, ++it) {
                const long long m = id / args.n_tiles;
                if (m != cur_m) {
                    mbar_wait(&sm.a_full, a_use & 1u);
                    ++a_use;
                    cur_m = m;
                }
                const uint32_t buf = it & 1u, duse = it >> 1;
                mbar_wait(&sm.d_empty[buf], (duse & 1u) ^ 1u);          // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem + buf * kFN;
                for (int kb = 0; kb < 4; ++kb, ++kbit) {
                    const int stage = kbit % kFStages;
                    const uint32_t use = kbit / kFStages;
                    mbar_wait(&sm.b_full[stage], use & 1u);
                    tc_fence_after();
                    const uint64_t a_hi = umma_smem_desc(&sm.a[0][kb][0]), a_lo = umma_smem_desc(&sm.a[1][kb][0]);
                    const uint64_t b_hi = umma_smem_desc(&sm.b[stage][0][0]), b_lo = umma_smem_desc(&sm.b[stage][1][0]);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {                    // 8 tf32 = 32 bytes = 2 descriptor units per step
                        const uint64_t o = (uint64_t)(ks * 2);
                        umma_tf32(d_tmem, a_lo + o, b_hi + o, idesc, (kb | ks) != 0);
                        umma_tf32(d_tmem, a_hi + o, b_lo + o, idesc, 1u);
                        umma_tf32(d_tmem, a_hi + o, b_hi + o, idesc, 1u);
                    }
                    umma_commit(&sm.b_empty[stage]);                    // frees the stage once these MMAs have read it
                }
                umma_commit(&sm.d_full[buf]);
                if (id + 1 == id_end || (id + 1) / args.n_tiles != m) umma_commit(&sm.a_empty);
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== epilogue: one query row per thread =====================
        const int ew = warp & 3, row = ew * 32 + lane;
        long long cur_m = -1, slot = -1, self_local = -1;
        float st = 0.0f;
        bool valid_q = false;
        int cgt = 0, cge = 0;
        uint32_t it = 0;
        auto flush = [&]() {
            if (valid_q && (cgt | cge)) {
                atomicAdd(&args.gt[slot], cgt);
                atomicAdd(&args.ge[slot], cge);
            }
            cgt = cge = 0;
        };
        for (long long id = id_begin; id < id_end; ++id, ++it) {
            const long long m = id / args.n_tiles, n = id % args.n_tiles;
            if (m != cur_m) {
                flush();
                const long long q = m * kFM + row;
                valid_q = q < 2 * args.b;
                if (valid_q) {
                    slot = q < args.b ? q : args.tail_off + (q - args.b);
                    st = args.true_score[slot];
                    self_local = args.self_id[q] - args.ent_offset;
                    valid_q = st == st;                                  // NaN = flagged triple (bad index)
                }
                cur_m = m;
            }
            const uint32_t buf = it & 1u, duse = it >> 1;
            mbar_wait(&sm.d_full[buf], duse & 1u);
            tc_fence_after();
            const long long tile_base = n * kFN;
            const int nvalid = (int)min((long long)kFN, args.n_local - tile_base);
            const long long self_col = self_local - tile_base;           // column of the true entity, if in this tile
            const uint32_t taddr = tmem + ((uint32_t)(ew * 32) << 16) + buf * kFN;
#pragma unroll 1
            for (int ch = 0; ch < kFN / 32; ++ch) {
                uint32_t v[32];
                __syncwarp();                                            // the TMEM load is warp-collective
                tmem_ld32(taddr + ch * 32, v);
                tmem_ld_wait();
                if (args.scores_out && m * kFM + row < 2 * args.b) {
                    float *orow = args.scores_out + (m * kFM + row) * args.ld_scores + tile_base + ch * 32;
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                        if (ch * 32 + c < nvalid) orow[c] = __uint_as_float(v[c]);
                }
                if (valid_q) {
                    if (nvalid == kFN) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            const float s = __uint_as_float(v[c]);
                            cgt += s > st;
                            cge += s >= st;
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            const float s = __uint_as_float(v[c]);
                            const bool ok = ch * 32 + c < nvalid;
                            cgt += ok && s > st;
                            cge += ok && s >= st;
                        }
                    }
                    if (self_col >= ch * 32 && self_col < ch * 32 + 32 && self_col < nvalid) {
                        // the true entity itself: it ties with s_true by definition (utils.py:104-105), whatever
                        // the fast arithmetic produced for it
                        const float s_self = __uint_as_float(select32(v, (int)(self_col & 31)));
                        cgt -= s_self > st;
                        cge += 1 - (s_self >= st);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.d_empty[buf]);
        }
        flush();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem, kTmemCols);
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
// [rows][128 floats] viewed as boxes of [128 rows][32 floats], 128-byte swizzle
static bool make_rows_tmap(CUtensorMap *tm, const void *base, long long rows) {
    EncodeTiledFn enc = fast_encode_fn();
    if (!enc) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)kD, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)kD * 4};
    const cuuint32_t box[2] = {32, 128};
    const cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }

long long fast_table_ws_bytes(long long n_local) { return 2 * round_up(n_local > 0 ? n_local : 1, kFN) * kD * 4; }
long long fast_query_ws_bytes(long long t) {
    const long long q_pad = round_up(2 * (t > 0 ? t : 1), kFM);
    return 2 * q_pad * kD * 4 + q_pad * 8;
}

int fast_prepare_table(const float *ent, long long n_local, void *table_ws, cudaStream_t st) {
    const long long n_pad = round_up(n_local > 0 ? n_local : 1, kFN);
    const long long total = n_pad * (kD / 4);
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    split_table_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float4 *>(ent), n_local, n_pad,
                                                         reinterpret_cast<float4 *>(table_ws));
    count_launch();
    return check_cuda(cudaGetLastError(), "split_table_kernel launch");
}

static int num_sms_fast() {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
}

// Folds the queries, then counts gt / ge over this shard with the tcgen05 kernel.  gt / ge / true_score must
// already hold zeros / the exact true scores (true_score_kernel has run on the same stream).
int launch_fast_sweep(int model, long long n_local, long long ent_offset, const RowRef &h, const RowRef &t, const RowRef &r,
                      const long long *triples, long long b, long long tail_off, const float *true_score, int *gt, int *ge,
                      const void *table_ws, void *query_ws, float *scores_out, long long ld_scores, cudaStream_t st) {
    if (model != BLP_MODEL_DISTMULT && model != BLP_MODEL_COMPLEX && model != BLP_MODEL_SIMPLE) {
        set_error("fast (tensor-core) mode covers distmult / complex / simple; transe is an L1 distance, not a contraction");
        return BLP_EINVAL;
    }
    FastArgs a{};
    a.n_local = n_local; a.ent_offset = ent_offset; a.n_pad = round_up(n_local, kFN);
    a.b = b; a.tail_off = tail_off; a.q_pad = round_up(2 * b, kFM);
    a.m_tiles = a.q_pad / kFM; a.n_tiles = a.n_pad / kFN;
    float *qsplit = reinterpret_cast<float *>(query_ws);
    long long *self_id = reinterpret_cast<long long *>(qsplit + 2 * a.q_pad * kD);
    a.true_score = true_score; a.self_id = self_id; a.gt = gt; a.ge = ge;
    a.scores_out = scores_out; a.ld_scores = ld_scores;

    switch (model) {
    case BLP_MODEL_DISTMULT: fold_queries_kernel<BLP_MODEL_DISTMULT><<<(unsigned)a.q_pad, kD, 0, st>>>(h, t, r, triples, b, a.q_pad, qsplit, self_id); break;
    case BLP_MODEL_COMPLEX: fold_queries_kernel<BLP_MODEL_COMPLEX><<<(unsigned)a.q_pad, kD, 0, st>>>(h, t, r, triples, b, a.q_pad, qsplit, self_id); break;
    default: fold_queries_kernel<BLP_MODEL_SIMPLE><<<(unsigned)a.q_pad, kD, 0, st>>>(h, t, r, triples, b, a.q_pad, qsplit, self_id); break;
    }
    count_launch();
    BLP_CUDA(cudaGetLastError());

    CUtensorMap tm_q, tm_e;
    memset(&tm_q, 0, sizeof(tm_q));
    memset(&tm_e, 0, sizeof(tm_e));
    if (!make_rows_tmap(&tm_q, qsplit, 2 * a.q_pad) || !make_rows_tmap(&tm_e, table_ws, 2 * a.n_pad)) {
        set_error("cuTensorMapEncodeTiled failed for the fast-mode operand tables");
        return BLP_ECUDA;
    }
    const size_t smem = sizeof(FastSmem) + 1024;
    BLP_CUDA(cudaFuncSetAttribute(fast_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long items = a.m_tiles * a.n_tiles;
    if (items == 0) return BLP_OK;
    const long long sms = num_sms_fast();
    const unsigned grid = (unsigned)(items < sms ? items : sms);
    prof_begin(1, st);
    fast_sweep_kernel<<<grid, kFThreads, smem, st>>>(a, tm_q, tm_e);
    prof_end(1, st);
    count_launch();
    BLP_CUDA(cudaGetLastError());
    return BLP_OK;
}

}  // namespace blp
