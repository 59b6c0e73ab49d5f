// blp_common.cuh -- shared device/host helpers for the BLP B200 kernels.
//
// Exact-arithmetic contract: every score computed by the eval path must carry
// the same fp32 roundings as the reference's CPU path:
// separate mul / add roundings (no FMA contraction), torch.norm(p=1) summed
// strictly sequentially, torch.sum(dim=-1) in ATen's 8-lane x 4-accumulator
// order.  All exact code uses the __f*_rn intrinsics, which nvcc never fuses.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/blp_b200.h"

namespace blp {

// ---- thread-local error / launch accounting (host) -------------------------
void set_error(const char *fmt, ...);
void count_launch(int n = 1);
void reset_launch_count();
int check_cuda(cudaError_t e, const char *what);
// measurement aid (blp_profile_events): bracket the selected kernel with the caller's CUDA events
void prof_begin(int which, cudaStream_t st);
void prof_end(int which, cudaStream_t st);

#define BLP_CUDA(call)                                 \
    do {                                               \
        int _rc = ::blp::check_cuda((call), #call);    \
        if (_rc) return _rc;                           \
    } while (0)

// ---- exact fp32 primitives --------------------------------------------------
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }

// ---- packed fp32 pairs (Blackwell FADD2 / FMUL2) -----------------------------
// One 64-bit register pair holds the same quantity for TWO queries; add / sub /
// mul round each half exactly like the scalar __f*_rn forms (no contraction), so
// packing changes no bits.  A packed instruction takes one issue slot for two
// lane-operations, which moves the exact-order sweeps from issue-bound to
// FP32-pipe-bound.  ptxas folds abs2 / dup2 into operand modifiers (|R|, R.F32).
typedef unsigned long long f2;
__device__ __forceinline__ f2 pack2(float lo, float hi) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f2 v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f2 dup2(float a) { return pack2(a, a); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
    f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {
    f2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// Packed multiply, rounded once per half.  ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even
// with --fmad=false (CUDA 12.9), which would change bits, so the product is issued as fma(a, b, -0):
// a*b + (-0) rounds to exactly fl(a*b) (including the sign of zero products), costs the same FFMA2 slot,
// and an fma result cannot be folded into a following add.  `nz` must be the pair {-0.0f, -0.0f} taken
// from a kernel argument (kNegZero2) so that ptxas cannot see its value and simplify the addend away.
constexpr f2 kNegZero2 = 0x8000000080000000ull;
__device__ __forceinline__ f2 mul2(f2 a, f2 b, f2 nz) {
    f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(nz));
    return r;
}
__device__ __forceinline__ f2 abs2(f2 a) {
    float lo, hi;
    unpack2(a, lo, hi);
    return pack2(fabsf(lo), fabsf(hi));
}

// ---- one product term of the bilinear models, natural row layout ------------
// models.py:226-227 / :230-239 / :242-248; j indexes [0, L), L = d (distmult)
// or d/2 (complex, simple).
template <int MODEL>
__device__ __forceinline__ float bilinear_term(const float *__restrict__ h, const float *__restrict__ t,
                                               const float *__restrict__ r, int j, int L) {
    if (MODEL == BLP_MODEL_DISTMULT) {
        return fmul(fmul(h[j], r[j]), t[j]);
    } else if (MODEL == BLP_MODEL_COMPLEX) {
        const float hr = h[j], hi = h[L + j], tr = t[j], ti = t[L + j], rr = r[j], ri = r[L + j];
        float p = fadd(fmul(fmul(rr, hr), tr), fmul(fmul(rr, hi), ti));
        p = fadd(p, fmul(fmul(ri, hr), ti));
        return fsub(p, fmul(fmul(ri, hi), tr));
    } else {
        const float hh = h[j], ht = h[L + j], th = t[j], tt = t[L + j], ra = r[j], rb = r[L + j];
        return fadd(fmul(fmul(hh, ra), tt), fmul(fmul(th, rb), ht));
    }
}

__device__ __forceinline__ int ceil_log2_i(int x) {
    if (x <= 2) return 1;
    return 32 - __clz(x - 1);
}

// ATen's CPU float row sum (SumKernel.cpp: vectorized_inner_sum -> row_sum ->
// multi_row_sum), any L.  term(j) yields element j.  Evaluated lane by lane so
// the state is 16 floats; the association is identical to the vector code.
template <class Term>
__device__ float aten_sum_generic(Term term, int L) {
    const int vec_size = L >> 3;
    const int size_ilp = vec_size >> 2;
    int level_power = ceil_log2_i(size_ilp) / 4;
    if (level_power < 4) level_power = 4;
    const int level_step = 1 << level_power;
    const int level_mask = level_step - 1;

    float fin = 0.0f;
    for (int j = vec_size * 8; j < L; ++j) fin = fadd(fin, term(j));
    for (int l = 0; l < 8; ++l) {
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[a][k] = 0.0f;
        int i = 0;
        while (i + level_step <= size_ilp) {
            for (int jj = 0; jj < level_step; ++jj, ++i) {
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[0][k] = fadd(acc[0][k], term(i * 32 + k * 8 + l));
            }
#pragma unroll
            for (int j = 1; j < 4; ++j) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    acc[j][k] = fadd(acc[j][k], acc[j - 1][k]);
                    acc[j - 1][k] = 0.0f;
                }
                if ((i & (level_mask << (j * level_power))) != 0) break;
            }
        }
        for (; i < size_ilp; ++i) {
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[0][k] = fadd(acc[0][k], term(i * 32 + k * 8 + l));
        }
#pragma unroll
        for (int j = 1; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[0][k] = fadd(acc[0][k], acc[j][k]);
        float a0 = acc[0][0];
        for (int v = size_ilp * 4; v < vec_size; ++v) a0 = fadd(a0, term(v * 8 + l));
#pragma unroll
        for (int k = 1; k < 4; ++k) a0 = fadd(a0, acc[0][k]);
        fin = fadd(fin, a0);
    }
    return fin;
}

// Exact score of one triple from natural-layout rows, any d (generic path;
// the d=128 sweep kernel in blp_eval.cu produces the same bits faster).
template <int MODEL>
__device__ float score_exact(const float *__restrict__ h, const float *__restrict__ t,
                             const float *__restrict__ r, int d) {
    if (MODEL == BLP_MODEL_TRANSE) {
        // models.py:222-223: x = fl(fl(h + r) - t); s = fl(s + |x|) sequentially
        float s = 0.0f;
        for (int j = 0; j < d; ++j) s = fadd(s, fabsf(fsub(fadd(h[j], r[j]), t[j])));
        return -s;
    } else if (MODEL == BLP_MODEL_DISTMULT) {
        return aten_sum_generic([&](int j) { return bilinear_term<MODEL>(h, t, r, j, d); }, d);
    } else {
        const int L = d >> 1;
        const float s = aten_sum_generic([&](int j) { return bilinear_term<MODEL>(h, t, r, j, L); }, L);
        return MODEL == BLP_MODEL_SIMPLE ? fmul(s, 0.5f) : s;   // models.py:247 "/ 2" is exact
    }
}

__device__ __forceinline__ float score_exact_dyn(int model, const float *__restrict__ h, const float *__restrict__ t,
                                                 const float *__restrict__ r, int d) {
    switch (model) {
    case BLP_MODEL_TRANSE: return score_exact<BLP_MODEL_TRANSE>(h, t, r, d);
    case BLP_MODEL_DISTMULT: return score_exact<BLP_MODEL_DISTMULT>(h, t, r, d);
    case BLP_MODEL_COMPLEX: return score_exact<BLP_MODEL_COMPLEX>(h, t, r, d);
    default: return score_exact<BLP_MODEL_SIMPLE>(h, t, r, d);
    }
}

// ---- mbarrier / bulk-copy (TMA) PTX wrappers --------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP), completion on an mbarrier.
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// 2-D tiled tensor copy global -> shared (TMA, SASS UTMALDG): box origin (c0 = innermost coordinate, c1 = row)
__device__ __forceinline__ void tma_tensor2d_g2s(void *dst_smem, const void *tmap, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst_smem)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

}  // namespace blp
