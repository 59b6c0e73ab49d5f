// blp_sampler.cu -- in-batch negative sampling indices on the device (a14, next row f3).
//
// Restates data.get_negative_sampling_indices (data.py:35-81): the 2B entities of a batch of B positive pairs
// are numbered row-major ([[0,1],[2,3],...]); every negative keeps one entity of its own pair and replaces
// the other -- head or tail chosen by a fair coin (data.py:71) -- with an entity drawn uniformly from the
// 2B - 2 entities of the OTHER rows (data.py:60-65: multinomial over weights that are zero on the own row).
// The reference draws on the CPU with torch's Mersenne-Twister-based multinomial, one (B, 2B) weight matrix
// per step, and ships B*K*2 int64 to the device; here one kernel writes the indices where the loss kernel
// reads them.  The random STREAM cannot match torch's (different generator), so parity is distributional:
// same support, same marginals, same memory layout / strides as the reference's return value.
//
// Layout: the reference returns a transposed view of a (K, B*repeats, 2) buffer (data.py:77-79), i.e. shape
// (B*repeats, K, 2) with element strides (2, 2*B*repeats, 1).  `out` is that buffer; with repeats > 1
// (DataParallel, data.py:297-298) column block c*B .. (c+1)*B belongs to device c and indexes ITS sub-batch.
//
// RNG: Philox4x32-10 keyed by the caller's seed, counter = (element index, stream offset): stateless,
// reproducible for a given (seed, offset), independent across elements.
#include "blp_common.cuh"

namespace blp {

struct Philox {
    uint32_t c[4];
};
__device__ __forceinline__ Philox philox4x32_10(uint64_t ctr_lo, uint64_t ctr_hi, uint64_t key) {
    uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
    uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return Philox{{c0, c1, c2, c3}};
}

template <typename IDX>
__global__ void negative_sample_kernel(long long batch, long long num_neg, long long repeats, uint64_t seed, uint64_t offset,
                                       IDX *__restrict__ out) {
    const long long cols = batch * repeats, total = num_neg * cols;
    const unsigned long long others = (unsigned long long)(2 * batch - 2);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = (i % cols) % batch;                       // pair inside its device's sub-batch
        const Philox r = philox4x32_10((uint64_t)i, offset, seed);
        // uniform over the 2B - 2 entities of the other rows: multiply-high of a 64-bit draw (bias < 2^-40)
        const unsigned long long u64 = ((unsigned long long)r.c[0] << 32) | r.c[1];
        long long repl = (long long)__umul64hi(u64, others);
        if (repl >= 2 * b) repl += 2;                                 // skip the own row (weights 0, data.py:62-63)
        const bool corrupt_tail = r.c[2] >> 31;                       // data.py:71 col_selector
        out[2 * i + 0] = (IDX)(corrupt_tail ? 2 * b : repl);
        out[2 * i + 1] = (IDX)(corrupt_tail ? repl : 2 * b + 1);
    }
}

}  // namespace blp

using namespace blp;

extern "C" int blp_negative_sample(int64_t batch, int64_t num_neg, int64_t repeats, uint64_t seed, uint64_t offset,
                                   int64_t *out, void *stream) {
    reset_launch_count();
    if (batch < 2) { set_error("negative sampling needs a batch of at least 2 pairs (got %lld)", (long long)batch); return BLP_EINVAL; }
    if (num_neg < 0 || repeats < 1) { set_error("bad num_neg / repeats"); return BLP_EINVAL; }
    const long long total = num_neg * batch * repeats;
    if (total == 0) return BLP_OK;
    if (!out) { set_error("null pointer argument"); return BLP_EINVAL; }
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    negative_sample_kernel<long long><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(batch, num_neg, repeats, seed, offset,
                                                                                        (long long *)out);
    count_launch();
    BLP_CUDA(cudaGetLastError());
    return BLP_OK;
}
