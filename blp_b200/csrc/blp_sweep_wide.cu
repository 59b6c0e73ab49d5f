// blp_sweep_wide.cu -- the fused TransE sweep for any row width d (d % 4 == 0, d != 128).
//
// Every BOW / DKRL script of the reference runs TransE on encoder outputs of width 300 (glove-bow) or 768 (bert-bow)
// (utils.py:11-19, scripts/{bert,glove}-*.sh, scripts/test-umls.sh); train.py:141-157 is the same loop at those widths.
// torch.norm(p=1) sums strictly sequentially over the row, which is chunk friendly: the accumulators
// of a (queries x candidates) register tile simply stay in registers while the row streams through shared memory in
// chunks of 64 floats.
//
// Same structure as blp_sweep.cu: persistent grid, 8 consumer warps + 1 TMA producer warp, work list of
// (triple group, 128-row candidate tile) items, queries folded into <= 2 operand vectors held pair-interleaved in shared
// memory (one 64-bit register feeds a packed FADD2: two queries per issue slot), exact fp32 operation order, two int32
// counters per query leave the SM.  Differences: the row is padded with zeros to a multiple of 32 floats (TMA fills the
// out-of-bounds columns with zeros; |fl(fl(0 + 0) - 0)| adds +0, which leaves every partial sum unchanged), a tile is
// consumed in ceil(d / 64) stages, and the register tile shrinks with d so that the folded queries still fit in shared
// memory (d <= 384: 16 triples per group, d <= 768: 8, d <= 1536: 4, d <= 3072: 2).
#include <cuda.h>
#include <string.h>

#include <atomic>

#include "blp_sweep.h"

namespace blp {

namespace {

constexpr int kCW = 8;                        // consumer warps
constexpr int kCT = 128;                      // candidate rows per tile
constexpr int kThreads = (kCW + 1) * 32;
constexpr int kBlk = 32;                      // floats per swizzled column block (one 128-byte swizzle span)
constexpr int kChunkBlks = 2;                 // column blocks per stage: 128 rows x 64 floats = 32 KB
constexpr int kStages = 3;
constexpr int kStageFloats = kCT * kBlk * kChunkBlks;

template <int TQP_, int TC_>
struct WCfg {
    static constexpr int TQP = TQP_;          // query pairs per consumer thread
    static constexpr int TC = TC_;            // candidates per consumer thread
    static constexpr int RS = 4 / TC_;        // warps that share one slot (they split the 128 tile rows)
    static constexpr int NS = kCW / RS;       // slots per CTA
    static constexpr int SQ = 2 * TQP_;       // queries per slot
    static constexpr int NQ = NS * SQ;        // queries per CTA
    static constexpr bool kMixed = TQP_ >= 2; // a slot predicts heads with its first TQP queries, tails with the rest
    static constexpr int kTriplesPerSlot = kMixed ? TQP_ : SQ;
    static constexpr int kRoleSlots = kMixed ? NS : NS / 2;
    static constexpr int kTriplesPerGroup = kRoleSlots * kTriplesPerSlot;
    __device__ static __forceinline__ bool is_head(int slot, int qi) { return kMixed ? qi < TQP_ : slot < NS / 2; }
    __device__ static __forceinline__ int triple(int slot, int qi) {
        return kMixed ? slot * TQP_ + (qi % TQP_) : (slot % (NS / 2)) * SQ + qi;
    }
};

struct WideArgs {
    const float *ent;
    long long n_local;
    RowRef h, t, r, h2, t2, r2;
    long long b, tail_off, groups;
    const float *true_score;
    int *gt, *ge;
    int d, nblk;                               // row width, 32-float blocks of the zero-padded row
};

__device__ __forceinline__ size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// four consecutive positions of one operand vector of one query pair: [pos][half] interleaved
struct Q4 {
    f2 x, y, z, w;
};
__device__ __forceinline__ Q4 ldq4(const float *vec, int off) {
    const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(vec + off * 2);
    const ulonglong2 b = *reinterpret_cast<const ulonglong2 *>(vec + off * 2 + 4);
    Q4 r;
    r.x = a.x; r.y = a.y; r.z = b.x; r.w = b.y;
    return r;
}

__device__ __forceinline__ void consumer_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kCW * 32) : "memory"); }

template <class C>
__global__ void __launch_bounds__(kThreads, 1) sweep_wide_kernel(const WideArgs args, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int dpad = args.nblk * kBlk;
    float *ctile = reinterpret_cast<float *>(base);                                   // [kStages][kStageFloats]
    float *qv = ctile + kStages * kStageFloats;                                       // [NQ / 2 pairs][2 vectors][dpad][2]
    float *st_s = qv + (size_t)(C::NQ / 2) * 2 * dpad * 2;                           // [NQ]
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(st_s + C::NQ);                  // [kStages]
    uint64_t *empty_bar = full_bar + kStages;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kCW);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const long long ntiles = (args.n_local + kCT - 1) / kCT;
    const long long total = ntiles * args.groups;
    const long long id_begin = total * blockIdx.x / gridDim.x, id_end = total * (blockIdx.x + 1) / gridDim.x;
    const int nchunks = (args.nblk + kChunkBlks - 1) / kChunkBlks;

    if (warp == kCW) {
        // ===================== producer warp: one elected lane feeds the stage ring with tensor copies =====================
        if (lane == 0) {
            unsigned it = 0;
            for (long long id = id_begin; id < id_end; ++id) {
                const int row0 = (int)((id % ntiles) * kCT);
                for (int ch = 0; ch < nchunks; ++ch, ++it) {
                    const int buf = it % kStages;
                    mbar_wait(&empty_bar[buf], ((it / kStages) & 1u) ^ 1u);
                    const int nb = min(kChunkBlks, args.nblk - ch * kChunkBlks);
                    mbar_arrive_expect_tx(&full_bar[buf], (uint32_t)(nb * kCT * kBlk * 4));
                    for (int k = 0; k < nb; ++k)       // rows / columns past the end of the table arrive as zeros
                        tma_tensor2d_g2s(ctile + buf * kStageFloats + k * (kCT * kBlk), &tmap, (ch * kChunkBlks + k) * kBlk, row0,
                                         &full_bar[buf]);
                }
            }
        }
        return;
    }

    // ===================== consumer warps =====================
    const int slot = warp / C::RS, rowgrp = warp % C::RS;
    const int row_off = rowgrp * C::TC * 32 + lane;               // this lane's first row inside a tile
    const int xr = (row_off & 7) << 2;                            // 128-byte swizzle: chunk c of row r sits at c ^ (r & 7)
    const float *qslot = qv + (size_t)(slot * C::TQP) * 2 * dpad * 2;
    unsigned it = 0;
    long long id = id_begin;
    while (id < id_end) {
        const long long grp = id / ntiles;
        const long long seg_end = min(id_end, (grp + 1) * ntiles);
        const long long t0 = grp * C::kTriplesPerGroup;

        consumer_bar_sync();                      // everyone is done with the previous group's vectors
        // query-side folding: one warp per query row, 16-byte chunks across the lanes (coalesced), zero padding past d
        for (int ql = warp; ql < C::NQ; ql += kCW) {
            const int s_ = ql / C::SQ, qi = ql % C::SQ;
            const bool hp = C::is_head(s_, qi);
            const long long tr = t0 + C::triple(s_, qi);
            float *qp = qv + ((size_t)(ql >> 1) * 2 * dpad) * 2 + (ql & 1);
            const bool live = tr < args.b;
            const float *h = nullptr, *t = nullptr, *r = nullptr;
            if (live) {
                h = (hp ? args.h : args.h2).row(tr, args.d);
                t = (hp ? args.t : args.t2).row(tr, args.d);
                r = (hp ? args.r : args.r2).row(tr, args.d);
            }
            for (int c = lane; c < dpad / 4; c += 32) {
                float4 hv = make_float4(0.f, 0.f, 0.f, 0.f), tv = hv, rv = hv;
                if (live && 4 * c < args.d) {
                    hv = __ldg(reinterpret_cast<const float4 *>(h) + c);
                    tv = __ldg(reinterpret_cast<const float4 *>(t) + c);
                    rv = __ldg(reinterpret_cast<const float4 *>(r) + c);
                }
                const float ha[4] = {hv.x, hv.y, hv.z, hv.w}, ta[4] = {tv.x, tv.y, tv.z, tv.w}, ra[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int j = 4 * c + k;
                    // head prediction: v0 = r, v1 = t (candidate plays heads: fl(fl(e + r) - t));
                    // tail prediction: v0 = fl(h + r) (fl(fl(h + r) - e)), v1 unused
                    qp[(size_t)j * 2] = hp ? ra[k] : fadd(ha[k], ra[k]);
                    qp[((size_t)dpad + j) * 2] = hp ? ta[k] : 0.0f;
                }
            }
        }
        if (tid < C::NQ) {
            const int s_ = tid / C::SQ, qi = tid % C::SQ;
            const long long tr = t0 + C::triple(s_, qi);
            const long long qo = (C::is_head(s_, qi) ? 0 : args.tail_off) + tr;
            st_s[tid] = tr < args.b ? args.true_score[qo] : 0.0f;
        }
        consumer_bar_sync();

        float st[C::SQ];
        long long qo[C::SQ];
        bool any = false;
#pragma unroll
        for (int q = 0; q < C::SQ; ++q) {
            st[q] = st_s[slot * C::SQ + q];
            const long long tr = t0 + C::triple(slot, q);
            qo[q] = tr < args.b ? (C::is_head(slot, q) ? 0 : args.tail_off) + tr : -1;
            any |= qo[q] >= 0;
        }
        int cgt[C::SQ], cge[C::SQ];
#pragma unroll
        for (int q = 0; q < C::SQ; ++q) cgt[q] = cge[q] = 0;

        for (; id < seg_end; ++id) {
            const long long tile = id % ntiles;
            f2 s[C::TQP][C::TC];
#pragma unroll
            for (int q = 0; q < C::TQP; ++q)
#pragma unroll
                for (int i = 0; i < C::TC; ++i) s[q][i] = 0ull;
            for (int ch = 0; ch < nchunks; ++ch, ++it) {
                const int buf = it % kStages;
                mbar_wait(&full_bar[buf], (it / kStages) & 1u);
                if (any) {
                    const int nb = min(kChunkBlks, args.nblk - ch * kChunkBlks);
                    const float *lane_base = ctile + buf * kStageFloats + row_off * kBlk;
#pragma unroll 1
                    for (int cb = 0; cb < nb; ++cb) {
                        const int pos0 = (ch * kChunkBlks + cb) * kBlk;           // first row position of this block
#pragma unroll
                        for (int cc = 0; cc < 8; ++cc) {
                            float e[C::TC][4];
#pragma unroll
                            for (int i = 0; i < C::TC; ++i) {
                                const float4 v = *reinterpret_cast<const float4 *>(lane_base + cb * (kCT * kBlk) + i * (32 * kBlk) + ((cc << 2) ^ xr));
                                e[i][0] = v.x; e[i][1] = v.y; e[i][2] = v.z; e[i][3] = v.w;
                            }
#pragma unroll
                            for (int q = 0; q < C::TQP; ++q) {
                                const bool head_pair = C::kMixed ? (q < C::TQP / 2) : (slot < C::NS / 2);
                                const float *v0 = qslot + ((size_t)q * 2 * dpad) * 2;
                                const Q4 a = ldq4(v0, pos0 + 4 * cc);
                                const f2 av[4] = {a.x, a.y, a.z, a.w};
                                if (head_pair) {
                                    const Q4 bq = ldq4(v0 + (size_t)dpad * 2, pos0 + 4 * cc);
                                    const f2 bv[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
#pragma unroll
                                        for (int i = 0; i < C::TC; ++i)
                                            s[q][i] = add2(s[q][i], abs2(sub2(add2(dup2(e[i][k]), av[k]), bv[k])));
                                } else {
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
#pragma unroll
                                        for (int i = 0; i < C::TC; ++i) s[q][i] = add2(s[q][i], abs2(sub2(av[k], dup2(e[i][k]))));
                                }
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[buf]);
            }
            if (!any) continue;
            const long long cbase = tile * kCT + row_off;
#pragma unroll
            for (int i = 0; i < C::TC; ++i) {
                const bool valid = cbase + 32 * i < args.n_local;
#pragma unroll
                for (int q = 0; q < C::TQP; ++q) {
                    float lo, hi;
                    unpack2(s[q][i], lo, hi);
                    lo = -lo; hi = -hi;                                       // score = -||.||_1
                    cgt[2 * q] += (valid && lo > st[2 * q]) ? 1 : 0;
                    cge[2 * q] += (valid && lo >= st[2 * q]) ? 1 : 0;
                    cgt[2 * q + 1] += (valid && hi > st[2 * q + 1]) ? 1 : 0;
                    cge[2 * q + 1] += (valid && hi >= st[2 * q + 1]) ? 1 : 0;
                }
            }
        }
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn wide_encode_fn() {
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

template <class C>
int launch_wide_cfg(WideArgs a, cudaStream_t st) {
    const int dpad = a.nblk * kBlk;
    const size_t smem = 1024 + (size_t)kStages * kStageFloats * 4 + (size_t)C::NQ * dpad * 2 * 4 + C::NQ * 4 + 2 * kStages * 8 + 64;
    if (smem > 227 * 1024) { set_error("sweep_wide: d = %d does not fit this register tile", a.d); return BLP_EDIM; }
    static std::atomic<int> attr_bytes[64];        // per device: largest dynamic smem size opted in so far
    int dev = 0;
    cudaGetDevice(&dev);
    dev = (dev >= 0 && dev < 64) ? dev : 0;
    if (attr_bytes[dev].load(std::memory_order_relaxed) < (int)smem) {
        BLP_CUDA(cudaFuncSetAttribute(sweep_wide_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_bytes[dev].store((int)smem, std::memory_order_relaxed);
    }
    a.groups = (a.b + C::kTriplesPerGroup - 1) / C::kTriplesPerGroup;
    const long long ntiles = (a.n_local + kCT - 1) / kCT;
    const long long items = ntiles * a.groups;
    if (items == 0) return BLP_OK;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned grid = (unsigned)(items < sms ? items : sms);
    EncodeTiledFn enc = wide_encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return BLP_ECUDA; }
    CUtensorMap tmap;
    const cuuint64_t dims[2] = {(cuuint64_t)a.d, (cuuint64_t)a.n_local};
    const cuuint64_t strides[1] = {(cuuint64_t)a.d * 4};
    const cuuint32_t box[2] = {kBlk, (cuuint32_t)kCT};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(a.ent), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed for a [%lld, %d] table", a.n_local, a.d);
        return BLP_ECUDA;
    }
    prof_begin(1, st);
    sweep_wide_kernel<C><<<grid, kThreads, smem, st>>>(a, tmap);
    prof_end(1, st);
    count_launch();
    BLP_CUDA(cudaGetLastError());
    return BLP_OK;
}

}  // namespace

bool sweep_wide_supports(int model, int d, const float *ent) {
    return model == BLP_MODEL_TRANSE && d > 0 && (d & 3) == 0 && d <= 3072 && (reinterpret_cast<uintptr_t>(ent) & 15u) == 0;
}

// TransE counting sweep for d % 4 == 0 (rows 16-byte aligned): expects true_score filled and gt / ge zeroed.
int launch_sweep_wide(const SweepArgs &s, int d, cudaStream_t st) {
    WideArgs a{};
    a.ent = s.ent; a.n_local = s.n_local; a.h = s.h; a.t = s.t; a.r = s.r; a.h2 = s.h2; a.t2 = s.t2; a.r2 = s.r2;
    a.b = s.b; a.tail_off = s.tail_off; a.true_score = s.true_score; a.gt = s.gt; a.ge = s.ge;
    a.d = d; a.nblk = (d + kBlk - 1) / kBlk;
    // the folded queries take NQ * dpad * 8 bytes of shared memory next to the 96 KB stage ring
    if (d <= 384 && s.b > 8) return launch_wide_cfg<WCfg<4, 2>>(a, st);     // 16 triples per group
    if (d <= 768 && s.b > 4) return launch_wide_cfg<WCfg<4, 1>>(a, st);     //  8
    if (d <= 1536 && s.b > 2) return launch_wide_cfg<WCfg<2, 1>>(a, st);    //  4
    return launch_wide_cfg<WCfg<1, 1>>(a, st);                              //  2
}

}  // namespace blp
