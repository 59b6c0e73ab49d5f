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

// fp32 reduction (red.global.add.v4.f32) throughput into a small L2-resident table with the access pattern of the
// gradient scatter of blp_train.cu: a group of 8 lanes adds one whole 128-float row (4 x 16 bytes per lane, 512
// contiguous bytes per group and instruction round) to a pseudo-random row, nothing else.  Bounds the backward pass of
// compute_loss at large B x K, which is paced by the L2 atomic units.
__global__ void __launch_bounds__(512) atomic_probe_kernel(float *table, unsigned int rows, int iters) {
    const int lane = threadIdx.x & 31, gl = lane & 7;
    unsigned int state = ((blockIdx.x * blockDim.x + threadIdx.x) >> 3) * 2654435761u + 12345u;   // one stream per lane group
    for (int it = 0; it < iters; ++it) {
        state = state * 1664525u + 1013904223u;
        float *row = table + (size_t)((state >> 8) % rows) * 128;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float *p = row + 4 * (gl + 8 * c);
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(1.0f), "f"(0.5f), "f"(0.25f), "f"(2.0f) : "memory");
        }
    }
}

}  // namespace blp

using namespace blp;

extern "C" int blp_atomic_probe(float *table, int64_t rows, int iters, double *bytes_host, void *stream) {
    reset_launch_count();
    if (!table || rows <= 0 || rows > (1ll << 30) || iters <= 0) { set_error("bad probe arguments"); return BLP_EINVAL; }
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned blocks = (unsigned)(sms > 0 ? sms : 148);
    atomic_probe_kernel<<<blocks, 512, 0, (cudaStream_t)stream>>>(table, (unsigned int)rows, iters);
    count_launch();
    BLP_CUDA(cudaGetLastError());
    if (bytes_host) *bytes_host = (double)blocks * (512 / 8) * 512.0 * iters;      // groups x 512 bytes per iteration
    return BLP_OK;
}

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
