// blp_probe.cu -- FP32 pipe micro-benchmarks (measurement aid, see include/blp_b200.h).
//
// The exact-order sweeps are bound by FP32 lane-ops, not by HBM, whenever
// ~10 or more queries share one pass over the table.  The
// denominator of that bound is measured on the device the bench runs on with
// the same instruction mixes the sweeps issue.
#include "blp_common.cuh"

namespace blp {

template <int VARIANT>
__global__ void __launch_bounds__(256) pipe_probe_kernel(float *sink, int iters, f2 nz) {
    constexpr int kChains = 16;
    const float seed = (float)(threadIdx.x & 7) * 0.125f + (float)blockIdx.x * 1e-9f;
    float out = 0.f;
    if (VARIANT == 0) {
        float a[kChains];
#pragma unroll
        for (int i = 0; i < kChains; ++i) a[i] = seed + (float)i;
        const float c = seed + 1.0f;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < kChains; ++i) a[i] = fadd(a[i], c);
        }
#pragma unroll
        for (int i = 0; i < kChains; ++i) out += a[i];
    } else if (VARIANT == 1) {
        float a[kChains], u[kChains];
#pragma unroll
        for (int i = 0; i < kChains; ++i) { a[i] = 0.f; u[i] = seed + (float)i; }
        float e = seed;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < kChains; ++i) a[i] = fadd(a[i], fabsf(fsub(u[i], e)));
            e = __int_as_float(__float_as_int(e) + 1);   // loop-variant operand: nothing can be hoisted or shared across iterations
        }
#pragma unroll
        for (int i = 0; i < kChains; ++i) out += a[i];
    } else if (VARIANT == 2) {
        f2 a[kChains / 2];
#pragma unroll
        for (int i = 0; i < kChains / 2; ++i) a[i] = pack2(seed + (float)i, seed - (float)i);
        const f2 c = pack2(seed + 1.0f, seed + 2.0f);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < kChains / 2; ++i) a[i] = add2(a[i], c);
        }
#pragma unroll
        for (int i = 0; i < kChains / 2; ++i) { float lo, hi; unpack2(a[i], lo, hi); out += lo + hi; }
    } else if (VARIANT == 3) {
        f2 a[kChains / 2], u[kChains / 2];
#pragma unroll
        for (int i = 0; i < kChains / 2; ++i) { a[i] = 0ull; u[i] = pack2(seed + (float)i, seed - (float)i); }
        float e = seed;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < kChains / 2; ++i) a[i] = add2(a[i], abs2(sub2(u[i], dup2(e))));
            e = __int_as_float(__float_as_int(e) + 1);
        }
#pragma unroll
        for (int i = 0; i < kChains / 2; ++i) { float lo, hi; unpack2(a[i], lo, hi); out += lo + hi; }
    } else if (VARIANT == 5) {
        f2 a[kChains / 2], v[kChains / 2];
#pragma unroll
        for (int i = 0; i < kChains / 2; ++i) { a[i] = 0ull; v[i] = pack2(seed + (float)i, seed - (float)i); }
        float e = seed;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < kChains / 2; ++i) a[i] = add2(a[i], mul2(v[i], dup2(e), nz));
            e = __int_as_float(__float_as_int(e) + 1);
        }
#pragma unroll
        for (int i = 0; i < kChains / 2; ++i) { float lo, hi; unpack2(a[i], lo, hi); out += lo + hi; }
    } else {
        float a[kChains], v[kChains];
#pragma unroll
        for (int i = 0; i < kChains; ++i) { a[i] = 0.f; v[i] = seed + (float)i; }
        float e = seed;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < kChains; ++i) a[i] = fadd(a[i], fmul(v[i], e));
            e = __int_as_float(__float_as_int(e) + 1);
        }
#pragma unroll
        for (int i = 0; i < kChains; ++i) out += a[i];
    }
    sink[(long long)blockIdx.x * blockDim.x + threadIdx.x] = out;
}

}  // namespace blp

using namespace blp;

extern "C" int blp_pipe_probe(int variant, float *sink, int64_t n_threads, int iters, double *lane_ops_host,
                              void *stream) {
    reset_launch_count();
    if (!sink || n_threads < 256 || iters <= 0 || variant < 0 || variant > 5) {
        set_error("bad probe arguments (variant 0..5, n_threads >= 256, iters > 0)");
        return BLP_EINVAL;
    }
    const int threads = 256;
    const unsigned blocks = (unsigned)(n_threads / threads);
    cudaStream_t st = (cudaStream_t)stream;
    switch (variant) {
    case 0: pipe_probe_kernel<0><<<blocks, threads, 0, st>>>(sink, iters, kNegZero2); break;
    case 1: pipe_probe_kernel<1><<<blocks, threads, 0, st>>>(sink, iters, kNegZero2); break;
    case 2: pipe_probe_kernel<2><<<blocks, threads, 0, st>>>(sink, iters, kNegZero2); break;
    case 3: pipe_probe_kernel<3><<<blocks, threads, 0, st>>>(sink, iters, kNegZero2); break;
    case 4: pipe_probe_kernel<4><<<blocks, threads, 0, st>>>(sink, iters, kNegZero2); break;
    default: pipe_probe_kernel<5><<<blocks, threads, 0, st>>>(sink, iters, kNegZero2); break;
    }
    count_launch();
    BLP_CUDA(cudaGetLastError());
    const double per_iter[6] = {16.0, 32.0, 16.0, 32.0, 32.0, 32.0};
    if (lane_ops_host) *lane_ops_host = (double)blocks * threads * per_iter[variant] * iters;
    return BLP_OK;
}
