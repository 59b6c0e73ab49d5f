// blp_table.cu -- entity-table production glue (a13, next row f4).
//
// The reference encodes the candidate entities in batches and copies each batch into one dense table on one
// device: `batch_emb = model(...)` (which ends in F.normalize for TransE, models.py:38-43) followed by
// `ent_emb[idx:idx + bs] = batch_emb` (train.py:95-123).  For an entity-sharded sweep every rank owns a row
// block of the table, so the glue becomes: normalise the encoder's raw output rows and write the ones this
// rank owns straight into its shard -- the full (N, D) table never exists on any single device.
//
// Arithmetic: F.normalize(x, dim=-1) = x / max(||x||_2, 1e-12) with ATen's CPU vector-norm order (8 lanes, one
// accumulator per lane over consecutive 8-element blocks, separate mul / add roundings, lanes folded
// sequentially from lane 0, scalar tail), IEEE sqrt and division -- bit-equal to the reference's CPU path
// (oracle/np_oracle.py l2_normalize_rows, tests/golden/normalize.npz).
#include <stdlib.h>

#include "blp_common.cuh"

namespace blp {

constexpr int kStoreWarps = 8;

// one warp per source row
__global__ void __launch_bounds__(kStoreWarps * 32) store_rows_kernel(const float *__restrict__ emb, long long m, int d,
                                                                      int normalize, const long long *__restrict__ dst_rows,
                                                                      long long row0, float *__restrict__ shard,
                                                                      long long n_local, long long ent_offset) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long i = (long long)blockIdx.x * kStoreWarps + warp;
    if (i >= m) return;
    const long long local = (dst_rows ? dst_rows[i] : row0 + i) - ent_offset;
    if (local < 0 || local >= n_local) return;                   // another rank owns this row
    const float *x = emb + i * d;
    float *y = shard + local * d;
    float denom = 1.0f;
    if (normalize) {
        const int full = d - d % 8;
        float acc = 0.0f;
        if (lane < 8)
            for (int b = lane; b < full; b += 8) acc = fadd(acc, fmul(x[b], x[b]));
        float s = 0.0f;
#pragma unroll
        for (int l = 0; l < 8; ++l) {
            const float a = __shfl_sync(0xffffffffu, acc, l);
            s = (l == 0) ? a : fadd(s, a);
        }
        for (int j = full; j < d; ++j) s = fadd(s, fmul(x[j], x[j]));
        denom = fmaxf(__fsqrt_rn(s), 1e-12f);                    // norm.clamp_min(eps)
    }
    for (int j = lane; j < d; j += 32) y[j] = normalize ? __fdiv_rn(x[j], denom) : x[j];
}

// d % 4 == 0, d <= 256, 16-byte aligned rows: a warp takes FOUR rows at a time.  The rows move through registers
// with coalesced 128-bit accesses (lane holds float4 number lane + 32 * u of each row, all loads in flight
// together); the squares are parked in shared memory so that lane group g = lane / 8 can replay ATen's per-lane
// accumulation order for row g (8 chains per row, 4 rows in parallel), the group's first lane order folds the 8
// lane sums, and every lane divides what it holds.
constexpr int kStoreRows = 4;
template <int V4>
__global__ void __launch_bounds__(kStoreWarps * 32) store_rows_vec_kernel(const float *__restrict__ emb, long long m, int d,
                                                                          int normalize, const long long *__restrict__ dst_rows,
                                                                          long long row0, float *__restrict__ shard,
                                                                          long long n_local, long long ent_offset) {
    extern __shared__ float sq_smem[];                               // [kStoreWarps][kStoreRows][d]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long first = ((long long)blockIdx.x * kStoreWarps + warp) * kStoreRows;
    if (first >= m) return;
    const int nv = d >> 2;
    float4 v[kStoreRows][V4];
    long long local[kStoreRows];
#pragma unroll
    for (int g = 0; g < kStoreRows; ++g) {
        const long long i = first + g;
        local[g] = -1;
        if (i < m) {
            local[g] = (dst_rows ? dst_rows[i] : row0 + i) - ent_offset;
            if (local[g] >= n_local) local[g] = -1;
        }
        const float4 *x4 = reinterpret_cast<const float4 *>(emb + i * d);
#pragma unroll
        for (int u = 0; u < V4; ++u) {
            v[g][u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (local[g] >= 0 && lane + 32 * u < nv) v[g][u] = __ldg(x4 + lane + 32 * u);
        }
    }
    float den[kStoreRows] = {1.f, 1.f, 1.f, 1.f};
    if (normalize) {
        float *sq = sq_smem + (size_t)warp * kStoreRows * d;
#pragma unroll
        for (int g = 0; g < kStoreRows; ++g)
#pragma unroll
            for (int u = 0; u < V4; ++u)
                if (lane + 32 * u < nv)
                    reinterpret_cast<float4 *>(sq + g * d)[lane + 32 * u] =
                        make_float4(fmul(v[g][u].x, v[g][u].x), fmul(v[g][u].y, v[g][u].y), fmul(v[g][u].z, v[g][u].z),
                                    fmul(v[g][u].w, v[g][u].w));
        __syncwarp();
        const int grp = lane >> 3, l = lane & 7;
        const float *row = sq + grp * d;
        const int full = d - d % 8;
        float acc = 0.0f;
        for (int b = l; b < full; b += 8) acc = fadd(acc, row[b]);
        float s = __shfl_sync(0xffffffffu, acc, grp * 8);
#pragma unroll
        for (int j = 1; j < 8; ++j) s = fadd(s, __shfl_sync(0xffffffffu, acc, grp * 8 + j));
        for (int j = full; j < d; ++j) s = fadd(s, row[j]);
        const float mine = fmaxf(__fsqrt_rn(s), 1e-12f);
#pragma unroll
        for (int g = 0; g < kStoreRows; ++g) den[g] = __shfl_sync(0xffffffffu, mine, g * 8);
    }
#pragma unroll
    for (int g = 0; g < kStoreRows; ++g) {
        if (local[g] < 0) continue;
        float4 *y4 = reinterpret_cast<float4 *>(shard + local[g] * d);
#pragma unroll
        for (int u = 0; u < V4; ++u)
            if (lane + 32 * u < nv) {
                float4 o = v[g][u];
                if (normalize)
                    o = make_float4(__fdiv_rn(o.x, den[g]), __fdiv_rn(o.y, den[g]), __fdiv_rn(o.z, den[g]), __fdiv_rn(o.w, den[g]));
                y4[lane + 32 * u] = o;
            }
    }
}

// ---- d == 128, contiguous destination: persistent TMA pipeline -----------------------------------------
// The register-tile kernel above has its loads in flight only while a CTA is in its load phase (every CTA loads, then
// normalises, then stores): 4.7 TB/s at 4.8 M rows, 5.6 TB/s without the normalisation, against 6.6 TB/s for a plain
// copy.  Here the bytes move through the TMA engine on both sides: a producer thread keeps kStreamStages bulk loads
// (32 rows = 16 KB each) in flight per CTA regardless of what the consumers do, the 8 consumer warps normalise their 4
// rows of a tile IN PLACE in shared memory (same operations in the same order as above: bit-equal), and one thread
// hands the tile to a bulk store; the stage is released when that store has read it (cp.async.bulk.wait_group.read).
constexpr int kStreamRows = kStoreWarps * kStoreRows;            // 32 rows per tile
constexpr int kStreamStages = 4;
constexpr int kStreamThreads = (kStoreWarps + 1) * 32;           // 8 consumer warps + the producer warp
struct __align__(128) StreamSmem {
    float tile[kStreamStages][kStreamRows][128];
    uint64_t full[kStreamStages];
    uint64_t empty[kStreamStages];
};
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kStoreWarps * 32) : "memory"); }

__global__ void __launch_bounds__(kStreamThreads) store_rows_stream_kernel(const float *__restrict__ emb, long long m, int normalize,
                                                                           float *__restrict__ dst) {
    extern __shared__ unsigned char stream_raw[];
    StreamSmem &sm = *reinterpret_cast<StreamSmem *>(stream_raw + ((128u - (smem_u32(stream_raw) & 127u)) & 127u));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStreamStages; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], 1);
        }
        mbar_fence_init();
    }
    __syncthreads();
    const long long tiles = (m + kStreamRows - 1) / kStreamRows;
    if (warp == kStoreWarps) {
        // ===== producer: one thread keeps kStreamStages bulk loads in flight =====
        if (lane == 0) {
            unsigned int it = 0;
            for (long long t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
                const int buf = it % kStreamStages;
                mbar_wait(&sm.empty[buf], ((it / kStreamStages) & 1u) ^ 1u);
                const long long r0 = t * kStreamRows;
                const uint32_t bytes = (uint32_t)(min((long long)kStreamRows, m - r0) * 128 * 4);
                mbar_arrive_expect_tx(&sm.full[buf], bytes);
                tma_bulk_g2s(&sm.tile[buf][0][0], emb + r0 * 128, bytes, &sm.full[buf]);
            }
        }
        return;
    }
    // ===== consumers =====
    unsigned int it = 0;
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
        const int buf = it % kStreamStages;
        mbar_wait(&sm.full[buf], (it / kStreamStages) & 1u);
        if (normalize) {
            float *rows = &sm.tile[buf][warp * kStoreRows][0];       // this warp's four rows
            float4 v[kStoreRows];
#pragma unroll
            for (int g = 0; g < kStoreRows; ++g) v[g] = reinterpret_cast<const float4 *>(rows + g * 128)[lane];
            // ATen's order for row grp = lane / 8: 8 lane chains over consecutive 8-element blocks, folded from lane 0
            const int grp = lane >> 3, l = lane & 7;
            const float *row = rows + grp * 128;
            float acc = 0.0f;
#pragma unroll
            for (int b = 0; b < 128; b += 8) acc = fadd(acc, fmul(row[b + l], row[b + l]));
            float s = __shfl_sync(0xffffffffu, acc, grp * 8);
#pragma unroll
            for (int j = 1; j < 8; ++j) s = fadd(s, __shfl_sync(0xffffffffu, acc, grp * 8 + j));
            const float mine = fmaxf(__fsqrt_rn(s), 1e-12f);         // norm.clamp_min(eps)
            __syncwarp();                                            // every lane has read the rows before they change
#pragma unroll
            for (int g = 0; g < kStoreRows; ++g) {
                const float den = __shfl_sync(0xffffffffu, mine, g * 8);
                reinterpret_cast<float4 *>(rows + g * 128)[lane] =
                    make_float4(__fdiv_rn(v[g].x, den), __fdiv_rn(v[g].y, den), __fdiv_rn(v[g].z, den), __fdiv_rn(v[g].w, den));
            }
            fence_async_smem();                                      // generic-proxy writes -> visible to the bulk store
        }
        consumer_sync();
        if (threadIdx.x == 0) {
            const long long r0 = t * kStreamRows;
            const uint32_t bytes = (uint32_t)(min((long long)kStreamRows, m - r0) * 128 * 4);
            bulk_s2g(dst + r0 * 128, &sm.tile[buf][0][0], bytes);
            bulk_commit();
            if (it > 0) {
                bulk_wait_read<1>();                                 // the previous tile's store has read its stage
                mbar_arrive(&sm.empty[(it - 1) % kStreamStages]);
            }
        }
    }
    if (threadIdx.x == 0) bulk_wait_read<0>();                       // shared memory must outlive the last store's read
}

}  // namespace blp

using namespace blp;

extern "C" int blp_store_rows(const float *emb, int64_t m, int d, int normalize, const int64_t *dst_rows, int64_t row0,
                              float *ent_shard, int64_t n_local, int64_t ent_offset, void *stream) {
    reset_launch_count();
    if (m < 0 || d <= 0 || n_local < 0) { set_error("bad size argument"); return BLP_EINVAL; }
    if (m == 0 || n_local == 0) return BLP_OK;
    if (!emb || !ent_shard) { set_error("null pointer argument"); return BLP_EINVAL; }
    const long long blocks = (m + kStoreWarps - 1) / kStoreWarps;
    const bool vec = (d % 4 == 0) && d <= 256 &&
                     ((reinterpret_cast<uintptr_t>(emb) | reinterpret_cast<uintptr_t>(ent_shard)) & 15u) == 0;
    static const int stream_env = []() {                              // tuning aid, read once: BLP_STORE_STREAM=0 disables the TMA path
        const char *e = getenv("BLP_STORE_STREAM");
        return e ? atoi(e) : 1;
    }();
    if (vec && d == 128 && !dst_rows && stream_env && m >= 4096) {
        // contiguous source rows -> contiguous destination rows: clip to the rows this rank owns on the host, then stream
        long long lo = ent_offset - row0, hi = ent_offset + n_local - row0;
        if (lo < 0) lo = 0;
        if (hi > m) hi = m;
        if (lo >= hi) return BLP_OK;                                  // another rank owns all of these rows
        const long long mm = hi - lo, tiles = (mm + kStreamRows - 1) / kStreamRows;
        const size_t smem = sizeof(StreamSmem) + 128;
        static bool attr_set[64] = {};
        int dev = 0;
        BLP_CUDA(cudaGetDevice(&dev));
        if (dev >= 0 && dev < 64 && !attr_set[dev]) {
            BLP_CUDA(cudaFuncSetAttribute(store_rows_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set[dev] = true;
        }
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const long long grid = tiles < 3ll * sms ? tiles : 3ll * sms;  // 3 CTAs of 64 KB per SM
        store_rows_stream_kernel<<<(unsigned)grid, kStreamThreads, smem, (cudaStream_t)stream>>>(
            emb + lo * 128, mm, normalize, ent_shard + (row0 + lo - ent_offset) * 128);
        count_launch();
        BLP_CUDA(cudaGetLastError());
        return BLP_OK;
    }
    if (vec) {
        const size_t smem = normalize ? (size_t)kStoreWarps * kStoreRows * d * sizeof(float) : 0;
        const long long vblocks = (m + kStoreWarps * kStoreRows - 1) / (kStoreWarps * kStoreRows);
        if (d <= 128)
            store_rows_vec_kernel<1><<<(unsigned)vblocks, kStoreWarps * 32, smem, (cudaStream_t)stream>>>(
                emb, m, d, normalize, (const long long *)dst_rows, row0, ent_shard, n_local, ent_offset);
        else
            store_rows_vec_kernel<2><<<(unsigned)vblocks, kStoreWarps * 32, smem, (cudaStream_t)stream>>>(
                emb, m, d, normalize, (const long long *)dst_rows, row0, ent_shard, n_local, ent_offset);
    } else {
        store_rows_kernel<<<(unsigned)blocks, kStoreWarps * 32, 0, (cudaStream_t)stream>>>(
            emb, m, d, normalize, (const long long *)dst_rows, row0, ent_shard, n_local, ent_offset);
    }
    count_launch();
    BLP_CUDA(cudaGetLastError());
    return BLP_OK;
}
