// blp_loss.cu -- stand-alone loss_fn / l2_regularization entry points (a5-a7).
//
// The training step goes through the fused blp_train_loss (blp_train.cu); these
// kernels serve direct callers of the reference's module-level functions
// margin_loss / nll_loss (models.py:251-258) and l2_regularization
// (models.py:261-266) on already materialised score tensors.  They are tiny
// (B*K <= a few 100 K elements): one CTA, warps own rows, fp64 accumulation of
// the mean so the result does not depend on the launch shape.
#include "blp_common.cuh"

namespace blp {

constexpr int kLossThreads = 1024;

__device__ __forceinline__ float softplus1(float x) { return x > 20.f ? x : log1pf(expf(x)); }   // F.softplus, threshold 20
__device__ __forceinline__ float sigmoid1(float x) {
    if (x > 20.f) return 1.f;
    const float z = expf(x);
    return z / (z + 1.f);
}

__device__ __forceinline__ double block_sum(double v, double *scratch) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double tot = 0.0;
    if (warp == 0) {
        tot = lane < (blockDim.x >> 5) ? scratch[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    }
    return tot;   // valid in warp 0
}

__global__ void __launch_bounds__(kLossThreads) pair_loss_kernel(int loss, const float *__restrict__ pos,
                                                                 const float *__restrict__ neg, long long neg_stride,
                                                                 long long b, long long k, float *__restrict__ loss_out,
                                                                 float *__restrict__ grad_pos, float *__restrict__ grad_neg) {
    __shared__ double scratch[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const float inv_bk = 1.0f / (float)(b * k);
    double acc = 0.0;
    for (long long row = warp; row < b; row += nwarps) {
        const float p = pos[row];
        const float omp = __fsub_rn(1.0f, p);
        float wsum = 0.f;
        for (long long j = lane; j < k; j += 32) {
            const float s = neg[row * neg_stride + j];
            float g;
            if (loss == BLP_LOSS_MARGIN) {
                const float m = __fadd_rn(omp, s);                // models.py:252
                const bool keep = !(m < 0.f);                     // models.py:253 zeroes m < 0 only
                if (keep) acc += (double)m;
                g = keep ? inv_bk : 0.f;
                wsum += g;
            } else {
                acc += (double)softplus1(s);                      // models.py:258
                g = 0.5f * inv_bk * sigmoid1(s);
            }
            if (grad_neg) grad_neg[row * k + j] = g;
        }
        if (loss == BLP_LOSS_MARGIN) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
            if (grad_pos && lane == 0) grad_pos[row] = -wsum;
        } else if (lane == 0) {
            acc += (double)softplus1(-p) * (double)k;             // rescaled below: mean over b, not b*k
            if (grad_pos) grad_pos[row] = -0.5f * sigmoid1(-p) / (float)b;
        }
    }
    const double tot = block_sum(acc, scratch);
    if (threadIdx.x == 0) {
        const double mean = tot / ((double)b * (double)k);
        *loss_out = (float)(loss == BLP_LOSS_MARGIN ? mean : 0.5 * mean);
    }
}

// (mean(h^2) + mean(t^2) + mean(r^2)) / 3  (models.py:261-266)
__global__ void __launch_bounds__(kLossThreads) l2_reg_kernel(const float *__restrict__ h, long long nh,
                                                              const float *__restrict__ t, long long nt,
                                                              const float *__restrict__ r, long long nr,
                                                              float *__restrict__ out) {
    __shared__ double scratch[32];
    double res = 0.0;
    const float *ptr[3] = {h, t, r};
    const long long cnt[3] = {nh, nt, nr};
    for (int a = 0; a < 3; ++a) {
        double acc = 0.0;
        for (long long i = threadIdx.x; i < cnt[a]; i += blockDim.x) acc += (double)ptr[a][i] * (double)ptr[a][i];
        const double tot = block_sum(acc, scratch);
        if (threadIdx.x == 0) res += tot / (double)cnt[a];
    }
    if (threadIdx.x == 0) *out = (float)(res / 3.0);
}

__global__ void scale_kernel(float *y, const float *x, const float *__restrict__ g, long long n) {
    const float s = *g;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = __fmul_rn(x[i], s);
}

}  // namespace blp

using namespace blp;

extern "C" int blp_pair_loss(int loss, const float *pos, const float *neg, int64_t neg_row_stride, int64_t b, int64_t k,
                             float *loss_out, float *grad_pos, float *grad_neg, void *stream) {
    reset_launch_count();
    if (loss != BLP_LOSS_MARGIN && loss != BLP_LOSS_NLL) { set_error("unknown loss id %d", loss); return BLP_EINVAL; }
    if (b <= 0 || k <= 0 || neg_row_stride < k) { set_error("bad shape b=%lld k=%lld stride=%lld", (long long)b, (long long)k, (long long)neg_row_stride); return BLP_EINVAL; }
    if (!pos || !neg || !loss_out) { set_error("null pointer argument"); return BLP_EINVAL; }
    pair_loss_kernel<<<1, kLossThreads, 0, (cudaStream_t)stream>>>(loss, pos, neg, neg_row_stride, b, k, loss_out, grad_pos, grad_neg);
    count_launch();
    BLP_CUDA(cudaGetLastError());
    return BLP_OK;
}

extern "C" int blp_l2_regularization(const float *heads, int64_t n_heads, const float *tails, int64_t n_tails,
                                     const float *rels, int64_t n_rels, float *out, void *stream) {
    reset_launch_count();
    if (n_heads <= 0 || n_tails <= 0 || n_rels <= 0) { set_error("l2_regularization needs non-empty tensors"); return BLP_EINVAL; }
    if (!heads || !tails || !rels || !out) { set_error("null pointer argument"); return BLP_EINVAL; }
    l2_reg_kernel<<<1, kLossThreads, 0, (cudaStream_t)stream>>>(heads, n_heads, tails, n_tails, rels, n_rels, out);
    count_launch();
    BLP_CUDA(cudaGetLastError());
    return BLP_OK;
}

extern "C" int blp_scale(float *y, const float *x, const float *scale_dev, int64_t n, void *stream) {
    reset_launch_count();
    if (n < 0) { set_error("negative size"); return BLP_EINVAL; }
    if (n == 0) return BLP_OK;
    if (!y || !x || !scale_dev) { set_error("null pointer argument"); return BLP_EINVAL; }
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    scale_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(y, x, scale_dev, n);
    count_launch();
    BLP_CUDA(cudaGetLastError());
    return BLP_OK;
}
