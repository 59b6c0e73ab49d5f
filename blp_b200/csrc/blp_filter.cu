// blp_filter.cu -- filtered ranks without per-batch host work (a12, next rows f1 / f2).
//
// The reference builds, for every eval batch, two dense (B, N) bool masks with Python loops over the
// filtering graph's edges (utils.get_triple_filters, utils.py:46-83), ships them to the device
// (train.py:160-164), overwrites the masked scores and runs get_metrics a second time (train.py:165-167).
// Semantically a filtered candidate just drops out of both rank counters, and the true entity is never
// filtered (utils.py:71,78), so the filtered ranks are a SPARSE CORRECTION of the raw ones.
//
// Here the filtering graph becomes a device-resident lookup structure built ONCE per evaluation:
//   tails_of  every edge as the composite key ((rel * n_rows + head_row) * n_rows + tail_row), sorted
//   heads_of  every edge as                   ((rel * n_rows + tail_row) * n_rows + head_row), sorted
// (entity ids already mapped through ent2idx, utils.py:31-43; edges touching an entity without a table row
// are dropped like `ent_idx != -1`, utils.py:72-73,79-80).  The known tails of (head, rel) are then one
// contiguous run found by two binary searches; parallel edges of the MultiDiGraph are adjacent duplicates.
// The correction kernel walks that run with one warp per query, re-scores only those candidates with the
// same exact-order code as the sweep, and subtracts them from the raw counters.
//
// Also here: the by-position / by-category MRR breakdowns of train.py:173-188 (utils.py:114-168) as one
// deterministic fp64 reduction on the device instead of per-triple Python loops with .item() syncs.
#include <cub/device/device_radix_sort.cuh>

#include "blp_sweep.h"

namespace blp {

struct FilterLayout {
    long long e_pad;        // entries per array, padded to 32
    size_t off_tails, off_heads, off_tmp, off_cub, cub_bytes, total;
};

static int filter_layout(long long num_edges, FilterLayout *L) {
    const long long e = num_edges > 0 ? num_edges : 1;
    L->e_pad = (e + 31) / 32 * 32;
    size_t cub_bytes = 0;
    const cudaError_t err = cub::DeviceRadixSort::SortKeys(nullptr, cub_bytes, (const unsigned long long *)nullptr,
                                                           (unsigned long long *)nullptr, (int)L->e_pad, 0, 64);
    if (err != cudaSuccess) return check_cuda(err, "cub::DeviceRadixSort::SortKeys (size query)");
    L->cub_bytes = (cub_bytes + 255) / 256 * 256;
    const size_t arr = (size_t)L->e_pad * 8;
    L->off_tails = 0;
    L->off_heads = arr;
    L->off_tmp = 2 * arr;
    L->off_cub = 3 * arr;
    L->total = 3 * arr + L->cub_bytes + 256;
    return BLP_OK;
}

// composite keys of every edge, both directions; edges without a table row / with a bad relation get the
// sentinel `invalid` (= num_rel * n_rows^2, larger than every valid key) and sort to the end
__global__ void filter_compose_kernel(const long long *__restrict__ edges, long long num_edges, long long e_pad,
                                      const long long *__restrict__ ent2idx, long long n_ids, long long n_rows,
                                      long long num_rel, unsigned long long invalid,
                                      unsigned long long *__restrict__ tails_key, unsigned long long *__restrict__ heads_key) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < e_pad; i += (long long)gridDim.x * blockDim.x) {
        unsigned long long kt = invalid, kh = invalid;
        if (i < num_edges) {
            long long h = edges[3 * i], t = edges[3 * i + 1];
            const long long r = edges[3 * i + 2];
            if (ent2idx) {                                       // utils.py:72,79: ent_idx = ent2idx[t]
                h = (h >= 0 && h < n_ids) ? ent2idx[h] : -1;
                t = (t >= 0 && t < n_ids) ? ent2idx[t] : -1;
            }
            if (h >= 0 && h < n_rows && t >= 0 && t < n_rows && r >= 0 && r < num_rel) {
                kt = ((unsigned long long)r * n_rows + h) * n_rows + t;
                kh = ((unsigned long long)r * n_rows + t) * n_rows + h;
            }
        }
        tails_key[i] = kt;
        heads_key[i] = kh;
    }
}

__device__ __forceinline__ long long lower_bound_u64(const unsigned long long *__restrict__ a, long long n, unsigned long long key) {
    long long lo = 0, hi = n;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// One warp per query (heads' queries first, then tails', the reference's cat order): remove the filtered
// candidates living in this shard from the raw counters (train.py:159-167).
__global__ void filter_correct_indexed_kernel(int model, const float *__restrict__ ent, long long n_local, long long ent_offset,
                                              int d, const RowRef hr, const RowRef tr, const RowRef rr,
                                              const long long *__restrict__ triples, long long b, long long tail_off,
                                              const unsigned long long *__restrict__ tails_of,
                                              const unsigned long long *__restrict__ heads_of, long long e_pad,
                                              long long n_rows, long long num_rel, const float *__restrict__ true_score,
                                              const int *__restrict__ gt, const int *__restrict__ ge, int *__restrict__ gt_f,
                                              int *__restrict__ ge_f) {
    const long long q = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= 2 * b) return;
    const bool head_pred = q < b;
    const long long i = head_pred ? q : q - b;
    const long long o = head_pred ? i : tail_off + i;
    const float st = true_score[o];
    const long long head = triples[3 * i], tail = triples[3 * i + 1], rel = triples[3 * i + 2];
    int cg = 0, ce = 0;
    const long long fixed = head_pred ? tail : head;            // the entity that stays in the query
    const long long truth = head_pred ? head : tail;            // never filtered (utils.py:71,78)
    if (fixed >= 0 && fixed < n_rows && rel >= 0 && rel < num_rel && st == st) {
        const unsigned long long *arr = head_pred ? heads_of : tails_of;
        const unsigned long long key = ((unsigned long long)rel * n_rows + fixed) * n_rows;
        const long long lo = lower_bound_u64(arr, e_pad, key), hi = lower_bound_u64(arr, e_pad, key + n_rows);
        const float *h = hr.row(i, d), *t = tr.row(i, d), *r = rr.row(i, d);
        for (long long p = lo + lane; p < hi; p += 32) {
            const unsigned long long comp = __ldg(arr + p);
            if (p > lo && __ldg(arr + p - 1) == comp) continue; // parallel edges collapse (MultiDiGraph)
            const long long cand = (long long)(comp - key);
            if (cand == truth) continue;
            const long long row = cand - ent_offset;
            if (row < 0 || row >= n_local) continue;            // another shard counts it
            const float *e = ent + row * d;
            const float s = head_pred ? score_exact_dyn(model, e, t, r, d) : score_exact_dyn(model, h, e, r, d);
            cg += s > st;
            ce += s >= st;
        }
    }
    cg = __reduce_add_sync(0xffffffffu, cg);
    ce = __reduce_add_sync(0xffffffffu, ce);
    if (lane == 0) {
        // Exact mode: the raw counters and this correction use the same arithmetic, so 0 <= gt_f < ge_f always holds
        // where the true entity lives.  Tensor-core mode counts with split-FP16 scores but corrects with exact ones: a
        // filtered candidate inside the tolerance band around s_true can be judged differently by the two, so the
        // filtered counters are clamped to their valid range (gt_f >= 0; ge_f >= gt_f, and > gt_f on the shard that
        // holds the true entity, which always ties itself).
        const long long trow = truth - ent_offset;
        const int g = max(gt[o] - cg, 0);
        int e = max(ge[o] - ce, g);
        if (trow >= 0 && trow < n_local && st == st) e = max(e, g + 1);
        gt_f[o] = g;
        ge_f[o] = e;
    }
}

// The same correction from the reference's own dense mask (train.py:160-165: `pred[filter_mask] = pred.min() - 1.0`):
// one warp per query scans its mask row and re-scores only the masked candidates.  Queries are independent
// (head-prediction queries [0, n_hq), then tail-prediction queries).  A masked TRUE candidate (never produced by
// utils.get_triple_filters, utils.py:71,78) gets the reference's semantics too: its score becomes min - 1, so every
// unmasked candidate ranks above it and every candidate ties or beats it.
__global__ void filter_correct_mask_kernel(int model, const float *__restrict__ ent, long long n_local, int d,
                                           const RowRef hq_h, const RowRef hq_t, const RowRef hq_r, long long n_hq,
                                           const RowRef tq_h, const RowRef tq_t, const RowRef tq_r, long long n_tq,
                                           const long long *__restrict__ hq_true, const long long *__restrict__ tq_true,
                                           long long ent_offset, const unsigned char *__restrict__ mask, long long ld_mask,
                                           const float *__restrict__ true_score, const int *__restrict__ gt,
                                           const int *__restrict__ ge, int *__restrict__ gt_f, int *__restrict__ ge_f) {
    const long long q = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= n_hq + n_tq) return;
    const bool head_pred = q < n_hq;
    const long long i = head_pred ? q : q - n_hq;
    const float st = true_score[q];
    const float *h = (head_pred ? hq_h : tq_h).row(i, d), *t = (head_pred ? hq_t : tq_t).row(i, d),
                *r = (head_pred ? hq_r : tq_r).row(i, d);
    const long long truth = (head_pred ? hq_true[i] : tq_true[i]) - ent_offset;
    const unsigned char *row = mask + q * ld_mask;
    int cg = 0, ce = 0, nm = 0;
    for (long long c = lane; c < n_local; c += 32) {
        if (!row[c]) continue;
        ++nm;
        if (c == truth) continue;
        const float *e = ent + c * d;
        const float s = head_pred ? score_exact_dyn(model, e, t, r, d) : score_exact_dyn(model, h, e, r, d);
        cg += s > st;
        ce += s >= st;
    }
    cg = __reduce_add_sync(0xffffffffu, cg);
    ce = __reduce_add_sync(0xffffffffu, ce);
    nm = __reduce_add_sync(0xffffffffu, nm);
    if (lane == 0) {
        const bool truth_masked = truth >= 0 && truth < n_local && row[truth];
        if (truth_masked) {
            gt_f[q] = (int)(n_local - nm);
            ge_f[q] = (int)n_local;
        } else {
            gt_f[q] = gt[q] - cg;
            ge_f[q] = ge[q] - ce;
        }
    }
}

// train.py:173-188: out[0..3) mrr_by_position (both new, head new, tail new), out[3..6) counts,
// out[6..14) mrr_by_category [2][4] (head prediction row, tail prediction row), out[14..18) category counts.
__global__ void __launch_bounds__(1024) mrr_breakdown_kernel(const float *__restrict__ recip, long long t, long long tail_off,
                                                             const long long *__restrict__ triples,
                                                             const unsigned char *__restrict__ is_new, long long n_ids,
                                                             const long long *__restrict__ rel_cat, long long num_rel,
                                                             double *__restrict__ out) {
    __shared__ double scratch[32][18];
    double acc[18];
#pragma unroll
    for (int j = 0; j < 18; ++j) acc[j] = 0.0;
    for (long long i = threadIdx.x; i < t; i += blockDim.x) {
        const float rh = recip[i], rt = recip[tail_off + i];
        const long long h = triples[3 * i], tl = triples[3 * i + 1], r = triples[3 * i + 2];
        if (is_new) {                                           // utils.split_by_new_position, utils.py:114-147
            const double v = (double)fadd(rh, rt) / 2.0;
            const bool hn = h >= 0 && h < n_ids && is_new[h], tn = tl >= 0 && tl < n_ids && is_new[tl];
            const int slot = (hn && tn) ? 0 : hn ? 1 : tn ? 2 : -1;
#pragma unroll
            for (int s = 0; s < 3; ++s)
                if (slot == s) { acc[s] += v; acc[3 + s] += 1.0; }
        }
        if (rel_cat && r >= 0 && r < num_rel) {                 // utils.split_by_category, utils.py:150-168
            const long long c = rel_cat[r];
#pragma unroll
            for (int s = 0; s < 4; ++s)
                if (c == s) { acc[6 + s] += (double)rh; acc[10 + s] += (double)rt; acc[14 + s] += 1.0; }
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < 18; ++j) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
        if (lane == 0) scratch[warp][j] = acc[j];
    }
    __syncthreads();
    if (threadIdx.x < 18) {
        double tot = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += scratch[w][threadIdx.x];
        out[threadIdx.x] = tot;
    }
}

}  // namespace blp

using namespace blp;

extern "C" int64_t blp_filter_index_bytes(int64_t num_edges) {
    FilterLayout L;
    if (num_edges < 0 || num_edges >= (1ll << 31) - 64 || filter_layout(num_edges, &L)) return -1;
    return (int64_t)L.total;
}

extern "C" int blp_filter_index_build(const int64_t *edges, int64_t num_edges, const int64_t *ent2idx, int64_t n_ids,
                                      int64_t n_rows, int64_t num_rel, void *index_ws, int64_t index_bytes, void *stream) {
    reset_launch_count();
    if (num_edges < 0 || num_edges >= (1ll << 31) - 64) { set_error("num_edges out of range"); return BLP_EINVAL; }
    if (n_rows <= 0 || num_rel <= 0 || !index_ws || (num_edges > 0 && !edges) || (ent2idx && n_ids <= 0)) {
        set_error("bad argument");
        return BLP_EINVAL;
    }
    // the composite key (rel * n_rows + a) * n_rows + b and its sentinel must fit 63 bits
    const long double span = (long double)num_rel * (long double)n_rows * (long double)n_rows;
    if (span >= 9.0e18L) { set_error("num_rel * n_rows^2 does not fit a 64-bit filter key"); return BLP_EINVAL; }
    FilterLayout L;
    int rc = filter_layout(num_edges, &L);
    if (rc) return rc;
    if ((size_t)index_bytes < L.total) { set_error("index_ws too small: %lld < %lld bytes", (long long)index_bytes, (long long)L.total); return BLP_EINVAL; }
    if (reinterpret_cast<uintptr_t>(index_ws) & 255u) { set_error("index_ws must be 256-byte aligned"); return BLP_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char *ws = reinterpret_cast<unsigned char *>(index_ws);
    unsigned long long *tails_of = reinterpret_cast<unsigned long long *>(ws + L.off_tails);
    unsigned long long *heads_of = reinterpret_cast<unsigned long long *>(ws + L.off_heads);
    unsigned long long *tmp = reinterpret_cast<unsigned long long *>(ws + L.off_tmp);
    const unsigned long long invalid = (unsigned long long)num_rel * (unsigned long long)n_rows * (unsigned long long)n_rows;
    int end_bit = 1;
    while (end_bit < 64 && (invalid >> end_bit) != 0) ++end_bit;

    // heads' keys are composed into their final array and sorted through tmp; tails' keys the other way round
    long long blocks = (L.e_pad + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    filter_compose_kernel<<<(unsigned)blocks, 256, 0, st>>>((const long long *)edges, num_edges, L.e_pad,
                                                            (const long long *)ent2idx, n_ids, n_rows, num_rel, invalid,
                                                            tmp, heads_of);
    count_launch();
    BLP_CUDA(cudaGetLastError());
    size_t cub_bytes = L.cub_bytes;
    BLP_CUDA(cub::DeviceRadixSort::SortKeys(ws + L.off_cub, cub_bytes, tmp, tails_of, (int)L.e_pad, 0, end_bit, st));
    BLP_CUDA(cudaMemcpyAsync(tmp, heads_of, (size_t)L.e_pad * 8, cudaMemcpyDeviceToDevice, st));
    cub_bytes = L.cub_bytes;
    BLP_CUDA(cub::DeviceRadixSort::SortKeys(ws + L.off_cub, cub_bytes, tmp, heads_of, (int)L.e_pad, 0, end_bit, st));
    count_launch(2);
    return BLP_OK;
}

extern "C" int blp_filter_correct(int model, const float *ent, int64_t n_local, int64_t ent_offset, int d,
                                  const float *rel_weight, int64_t num_rel, const int64_t *triples, int64_t t,
                                  const float *h_rows, const float *t_rows, const void *index_ws, int64_t num_edges,
                                  int64_t n_rows, int64_t tail_off, const float *true_score, const int32_t *gt,
                                  const int32_t *ge, int32_t *gt_f, int32_t *ge_f, void *stream) {
    reset_launch_count();
    if (model < 0 || model > 3) { set_error("unknown relational model id %d", model); return BLP_EINVAL; }
    if (d <= 0 || ((model == BLP_MODEL_COMPLEX || model == BLP_MODEL_SIMPLE) && (d & 1))) { set_error("bad d %d", d); return BLP_EDIM; }
    if (t < 0 || n_local < 0 || num_edges < 0 || n_rows <= 0 || num_rel <= 0 || tail_off < t) { set_error("bad size argument"); return BLP_EINVAL; }
    if (t == 0) return BLP_OK;
    if (!rel_weight || !triples || !index_ws || !true_score || !gt || !ge || !gt_f || !ge_f || (n_local > 0 && !ent)) {
        set_error("null pointer argument");
        return BLP_EINVAL;
    }
    if ((h_rows == nullptr) != (t_rows == nullptr)) { set_error("h_rows and t_rows must both be given or both NULL"); return BLP_EINVAL; }
    if (!h_rows && n_local == 0) { set_error("cannot gather query rows from an empty shard; pass h_rows / t_rows"); return BLP_EINVAL; }
    FilterLayout L;
    int rc = filter_layout(num_edges, &L);
    if (rc) return rc;
    const long long *tr = (const long long *)triples;
    const RowRef h = h_rows ? dense_rows(h_rows) : RowRef{ent, tr + 0, 3, ent_offset, n_local};
    const RowRef tt = t_rows ? dense_rows(t_rows) : RowRef{ent, tr + 1, 3, ent_offset, n_local};
    const RowRef r = RowRef{rel_weight, tr + 2, 3, 0, num_rel};
    const unsigned char *ws = reinterpret_cast<const unsigned char *>(index_ws);
    const int threads = 128;
    const long long blocks = (2 * t * 32 + threads - 1) / threads;
    filter_correct_indexed_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        model, ent, n_local, ent_offset, d, h, tt, r, tr, t, tail_off,
        reinterpret_cast<const unsigned long long *>(ws + L.off_tails),
        reinterpret_cast<const unsigned long long *>(ws + L.off_heads), L.e_pad, n_rows, num_rel, true_score, gt, ge, gt_f, ge_f);
    count_launch();
    BLP_CUDA(cudaGetLastError());
    return BLP_OK;
}

// blp_filter_correct for the score-matrix path (train.py:141-171 untouched): the lookup keys come from
// `triples` (head row, tail row, relation id), the operand rows are the dense (t, d) blocks the call site gathered
// (head_embs, tail_embs, rel_embs of train.py:141-143).
extern "C" int blp_filter_correct_rows(int model, const float *ent, int64_t n_local, int64_t ent_offset, int d,
                                       const int64_t *triples, int64_t t, const float *h_rows, const float *t_rows,
                                       const float *r_rows, const void *index_ws, int64_t num_edges, int64_t n_rows,
                                       int64_t num_rel, int64_t tail_off, const float *true_score, const int32_t *gt,
                                       const int32_t *ge, int32_t *gt_f, int32_t *ge_f, void *stream) {
    reset_launch_count();
    if (model < 0 || model > 3) { set_error("unknown relational model id %d", model); return BLP_EINVAL; }
    if (d <= 0 || ((model == BLP_MODEL_COMPLEX || model == BLP_MODEL_SIMPLE) && (d & 1))) { set_error("bad d %d", d); return BLP_EDIM; }
    if (t < 0 || n_local < 0 || num_edges < 0 || n_rows <= 0 || num_rel <= 0 || tail_off < t) { set_error("bad size argument"); return BLP_EINVAL; }
    if (t == 0) return BLP_OK;
    if (!triples || !h_rows || !t_rows || !r_rows || !index_ws || !true_score || !gt || !ge || !gt_f || !ge_f || (n_local > 0 && !ent)) {
        set_error("null pointer argument");
        return BLP_EINVAL;
    }
    FilterLayout L;
    int rc = filter_layout(num_edges, &L);
    if (rc) return rc;
    const unsigned char *ws = reinterpret_cast<const unsigned char *>(index_ws);
    const int threads = 128;
    const long long blocks = (2 * t * 32 + threads - 1) / threads;
    filter_correct_indexed_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        model, ent, n_local, ent_offset, d, dense_rows(h_rows), dense_rows(t_rows), dense_rows(r_rows), (const long long *)triples, t,
        tail_off, reinterpret_cast<const unsigned long long *>(ws + L.off_tails),
        reinterpret_cast<const unsigned long long *>(ws + L.off_heads), L.e_pad, n_rows, num_rel, true_score, gt, ge, gt_f, ge_f);
    count_launch();
    BLP_CUDA(cudaGetLastError());
    return BLP_OK;
}

// train.py:164-167 from the reference's own dense bool mask: filtered counters of the n_hq + n_tq queries of
// blp_rank_queries (same query arguments, same output order).  mask is (n_hq + n_tq, ld_mask) bytes, column j =
// candidate row ent_offset + j.
extern "C" int blp_filter_correct_mask(int model, const float *ent, int64_t n_local, int64_t ent_offset, int d,
                                       const float *hq_tails, const float *hq_rels, const int64_t *hq_true, int64_t n_hq,
                                       const float *tq_heads, const float *tq_rels, const int64_t *tq_true, int64_t n_tq,
                                       const uint8_t *mask, int64_t ld_mask, const float *true_score, const int32_t *gt,
                                       const int32_t *ge, int32_t *gt_f, int32_t *ge_f, void *stream) {
    reset_launch_count();
    if (model < 0 || model > 3) { set_error("unknown relational model id %d", model); return BLP_EINVAL; }
    if (d <= 0 || ((model == BLP_MODEL_COMPLEX || model == BLP_MODEL_SIMPLE) && (d & 1))) { set_error("bad d %d", d); return BLP_EDIM; }
    if (n_hq < 0 || n_tq < 0 || n_local <= 0 || ld_mask < n_local) { set_error("bad size argument"); return BLP_EINVAL; }
    const int64_t nq = n_hq + n_tq;
    if (nq == 0) return BLP_OK;
    if (!ent || !mask || !true_score || !gt || !ge || !gt_f || !ge_f || (n_hq > 0 && (!hq_tails || !hq_rels || !hq_true)) ||
        (n_tq > 0 && (!tq_heads || !tq_rels || !tq_true))) {
        set_error("null pointer argument");
        return BLP_EINVAL;
    }
    const RowRef hq_h = RowRef{ent, (const long long *)hq_true, 1, ent_offset, n_local};
    const RowRef tq_t = RowRef{ent, (const long long *)tq_true, 1, ent_offset, n_local};
    const int threads = 128;
    const long long blocks = (nq * 32 + threads - 1) / threads;
    filter_correct_mask_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        model, ent, n_local, d, hq_h, dense_rows(hq_tails), dense_rows(hq_rels), n_hq, dense_rows(tq_heads), tq_t,
        dense_rows(tq_rels), n_tq, (const long long *)hq_true, (const long long *)tq_true, ent_offset, mask, ld_mask, true_score,
        gt, ge, gt_f, ge_f);
    count_launch();
    BLP_CUDA(cudaGetLastError());
    return BLP_OK;
}

extern "C" int blp_mrr_breakdown(const float *recip, int64_t t, int64_t tail_off, const int64_t *triples,
                                 const uint8_t *is_new, int64_t n_ids, const int64_t *rel_categories, int64_t num_rel,
                                 double *out, void *stream) {
    reset_launch_count();
    if (t < 0 || tail_off < t || !out || (t > 0 && (!recip || !triples))) { set_error("bad argument"); return BLP_EINVAL; }
    if ((is_new && n_ids <= 0) || (rel_categories && num_rel <= 0)) { set_error("bad table size"); return BLP_EINVAL; }
    mrr_breakdown_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(recip, t, tail_off, (const long long *)triples, is_new, n_ids,
                                                               (const long long *)rel_categories, num_rel, out);
    count_launch();
    BLP_CUDA(cudaGetLastError());
    return BLP_OK;
}
