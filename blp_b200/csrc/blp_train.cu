// blp_train.cu -- fused LinkPrediction.compute_loss forward + backward (a5-a8).
//
// Replaces models.py:51-70 (rel lookup, positive scores, in-batch negative
// gather `ent_embs.view(2B, D)[neg_idx]`, negative scores, margin / NLL loss,
// optional L2 regulariser) and the autograd graph behind it with ONE kernel:
// the (B, K, 2, D) gather, the (B, K, D) broadcast intermediates and the
// (B, K) score matrix of the reference are never materialised (neg_scores is
// written only because callers may ask for it).
//
// Mapping: CTA = (positive row b, slice of its K negatives); one warp scores a
// negative with the dim axis spread over its lanes (coalesced 8-byte vector
// loads of the two gathered rows, warp-shuffle reduction), keeps several
// negatives in flight to cover L2 latency, and -- because the loss is linear in
// the upstream gradient -- immediately accumulates d loss / d rows:
//   * the positive rows (2b, 2b+1) and the relation row accumulate in registers
//     and are flushed once per warp;
//   * rows sampled as corrupting entities take vector reductions (red.global)
//     into grad_ent.
// The working set (2B rows, the int64 index tensor) is L2 resident; at B=64 the
// whole step is launch-latency bound, which is why it is a single launch plus
// one memset.  The scalar loss is reduced deterministically by the last CTA.
#include <stdlib.h>

#include "blp_common.cuh"

namespace blp {

constexpr int kTrainThreads = 512;
constexpr int kTrainWarps = kTrainThreads / 32;

struct TrainArgs {
    const float *ent;          // [2b, d]
    const float *rel_weight;   // [num_rel, d]
    const long long *rels;     // [b]
    const long long *neg_idx;  // strided (b, k, 2)
    long long s0, s1, s2;
    long long b, k, num_rel;
    int d;
    int slices;                // CTAs per positive row
    int loss;
    float regularizer;
    float *loss_out;
    float *pos_scores;
    float *neg_scores;
    float *grad_ent;
    float *grad_rel;            // [num_rel, d], rows rels[b] accumulate
    unsigned int *counter;     // workspace[0]
    float *partials;           // workspace + 16 floats
    int *err_flag;             // workspace[1]: set when an index is out of range
};

template <int MODEL>
struct TM {
    static constexpr bool kHalves = (MODEL == BLP_MODEL_COMPLEX || MODEL == BLP_MODEL_SIMPLE);
};

template <int NCH2>
struct Row {
    float2 lo[NCH2];
    float2 hi[NCH2];
};

template <int MODEL, int NCH2>
__device__ __forceinline__ void load_row(Row<NCH2> &x, const float *__restrict__ row, int P, int lane) {
#pragma unroll
    for (int c = 0; c < NCH2; ++c) {
        const int p = 2 * lane + 64 * c;
        x.lo[c] = make_float2(0.f, 0.f);
        x.hi[c] = make_float2(0.f, 0.f);
        if (p < P) {
            x.lo[c] = *reinterpret_cast<const float2 *>(row + p);
            if (TM<MODEL>::kHalves) x.hi[c] = *reinterpret_cast<const float2 *>(row + P + p);
        }
    }
}

template <int MODEL>
__device__ __forceinline__ float term1(float hl, float hh, float tl, float th, float rl, float rh) {
    if (MODEL == BLP_MODEL_TRANSE) return fabsf((hl + rl) - tl);
    if (MODEL == BLP_MODEL_DISTMULT) return (hl * rl) * tl;
    if (MODEL == BLP_MODEL_COMPLEX) return rl * hl * tl + rl * hh * th + rh * hl * th - rh * hh * tl;
    return hl * rl * th + tl * rh * hh;
}

template <int MODEL, int NCH2>
__device__ __forceinline__ float partial_score(const Row<NCH2> &h, const Row<NCH2> &t, const Row<NCH2> &r) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < NCH2; ++c) {
        s += term1<MODEL>(h.lo[c].x, h.hi[c].x, t.lo[c].x, t.hi[c].x, r.lo[c].x, r.hi[c].x);
        s += term1<MODEL>(h.lo[c].y, h.hi[c].y, t.lo[c].y, t.hi[c].y, r.lo[c].y, r.hi[c].y);
    }
    return s;
}

template <int MODEL>
__device__ __forceinline__ float finish_score(float s) {
    if (MODEL == BLP_MODEL_TRANSE) return -s;
    if (MODEL == BLP_MODEL_SIMPLE) return 0.5f * s;
    return s;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// d score / d (h, t, r) at one position, scaled by w (SimplE's 1/2 folded into w by the caller)
template <int MODEL>
__device__ __forceinline__ void grad1(float w, float hl, float hh, float tl, float th, float rl, float rh,
                                      float &dhl, float &dhh, float &dtl, float &dth, float &drl, float &drh) {
    if (MODEL == BLP_MODEL_TRANSE) {
        const float x = (hl + rl) - tl;
        const float sg = (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f);   // sign(0) = 0, like torch
        dhl = -w * sg; drl = -w * sg; dtl = w * sg;
        dhh = dth = drh = 0.f;
    } else if (MODEL == BLP_MODEL_DISTMULT) {
        dhl = w * rl * tl; dtl = w * hl * rl; drl = w * hl * tl;
        dhh = dth = drh = 0.f;
    } else if (MODEL == BLP_MODEL_COMPLEX) {   // l = re, h = im
        dhl = w * (rl * tl + rh * th);
        dhh = w * (rl * th - rh * tl);
        dtl = w * (rl * hl - rh * hh);
        dth = w * (rl * hh + rh * hl);
        drl = w * (hl * tl + hh * th);
        drh = w * (hl * th - hh * tl);
    } else {   // SIMPLE: heads = (hh_, ht_) = (hl, hh); tails = (th_, tt_) = (tl, th); rels = (ra, rb) = (rl, rh)
        dhl = w * rl * th;
        dhh = w * tl * rh;
        dtl = w * rh * hh;
        dth = w * hl * rl;
        drl = w * hl * th;
        drh = w * tl * hh;
    }
}

__device__ __forceinline__ void red_add_v2(float *addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

template <int MODEL, int NCH2>
struct Acc {
    Row<NCH2> h, t, r;
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int c = 0; c < NCH2; ++c) {
            h.lo[c] = h.hi[c] = t.lo[c] = t.hi[c] = r.lo[c] = r.hi[c] = make_float2(0.f, 0.f);
        }
    }
};

template <int MODEL, int NCH2>
__device__ __forceinline__ void flush_row(float *__restrict__ dst, const Row<NCH2> &g, int P, int lane) {
#pragma unroll
    for (int c = 0; c < NCH2; ++c) {
        const int p = 2 * lane + 64 * c;
        if (p < P) {
            red_add_v2(dst + p, g.lo[c].x, g.lo[c].y);
            if (TM<MODEL>::kHalves) red_add_v2(dst + P + p, g.hi[c].x, g.hi[c].y);
        }
    }
}

// accumulate w * dscore into (gh, gt, gr); each may be a register accumulator
template <int MODEL, int NCH2>
__device__ __forceinline__ void accumulate(float w, const Row<NCH2> &h, const Row<NCH2> &t, const Row<NCH2> &r,
                                           Row<NCH2> &gh, Row<NCH2> &gt, Row<NCH2> &gr) {
#pragma unroll
    for (int c = 0; c < NCH2; ++c) {
        float a, b2, c2, d2, e, f;
        grad1<MODEL>(w, h.lo[c].x, h.hi[c].x, t.lo[c].x, t.hi[c].x, r.lo[c].x, r.hi[c].x, a, b2, c2, d2, e, f);
        gh.lo[c].x += a; gh.hi[c].x += b2; gt.lo[c].x += c2; gt.hi[c].x += d2; gr.lo[c].x += e; gr.hi[c].x += f;
        grad1<MODEL>(w, h.lo[c].y, h.hi[c].y, t.lo[c].y, t.hi[c].y, r.lo[c].y, r.hi[c].y, a, b2, c2, d2, e, f);
        gh.lo[c].y += a; gh.hi[c].y += b2; gt.lo[c].y += c2; gt.hi[c].y += d2; gr.lo[c].y += e; gr.hi[c].y += f;
    }
}

__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float softplus_grad_f(float x) {
    if (x > 20.f) return 1.f;
    const float z = expf(x);
    return z / (z + 1.f);
}

#ifndef BLP_TRAIN_MINB
#define BLP_TRAIN_MINB 1
#endif
template <int MODEL, int NCH2, int ILP, bool GRAD>
__global__ void __launch_bounds__(kTrainThreads, (NCH2 <= 2 ? BLP_TRAIN_MINB : 1)) train_kernel(const TrainArgs a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long b = blockIdx.x / a.slices;
    const int slice = blockIdx.x % a.slices;
    const int d = a.d;
    const int P = TM<MODEL>::kHalves ? d / 2 : d;          // positions owned pairwise by the lanes
    const float half = (MODEL == BLP_MODEL_SIMPLE) ? 0.5f : 1.0f;
    const long long nb2 = 2 * a.b;

    // positive triple of this row (models.py:55-57)
    long long rel = a.rels[b];
    if (rel < 0 || rel >= a.num_rel) {
        if (threadIdx.x == 0) *a.err_flag = 1;
        rel = 0;
    }
    Row<NCH2> hb, tb, rb;
    load_row<MODEL, NCH2>(hb, a.ent + (2 * b) * d, P, lane);
    load_row<MODEL, NCH2>(tb, a.ent + (2 * b + 1) * d, P, lane);
    load_row<MODEL, NCH2>(rb, a.rel_weight + rel * d, P, lane);
    const float pos = finish_score<MODEL>(warp_sum(partial_score<MODEL, NCH2>(hb, tb, rb)));
    const float one_minus_pos = 1.0f - pos;

    const float inv_bk = 1.0f / (float)(a.b * a.k);
    Acc<MODEL, NCH2> acc;
    if (GRAD) acc.zero();
    float loss_part = 0.f;    // lane-uniform
    float wsum = 0.f;         // sum of negative weights handled by this warp (margin: feeds d/dpos)

    const long long kps = (a.k + a.slices - 1) / a.slices;
    const long long kb = slice * kps;
    const long long ke = min(a.k, kb + kps);
    const long long *nrow = a.neg_idx + b * a.s0;

    // every warp owns a contiguous run of this slice's negatives; it fetches the index pairs of 32 negatives with one
    // load per lane (the indices of one row are 2B * 8 bytes apart in the reference sampler's layout, data.py:77-79)
    // and hands them out by shuffle, so the dependent index -> row load chain is paid once per 32 negatives
    const long long per_warp = (ke - kb + kTrainWarps - 1) / kTrainWarps;
    const long long wb = kb + (long long)warp * per_warp, we = min(ke, wb + per_warp);
    for (long long base = wb; base < we; base += 32) {
    long long li0 = 2 * b, li1 = 2 * b + 1;
    if (base + lane < we) {
        li0 = nrow[(base + lane) * a.s1];
        li1 = nrow[(base + lane) * a.s1 + a.s2];
        if (li0 < 0 || li0 >= nb2 || li1 < 0 || li1 >= nb2) {
            *a.err_flag = 1;
            li0 = 2 * b; li1 = 2 * b + 1;
        }
    }
    const int cnt = (int)min(32ll, we - base);
    for (int u0 = 0; u0 < cnt; u0 += ILP) {
        const long long k0 = base + u0;
        long long i0[ILP], i1[ILP];
        Row<NCH2> nh[ILP], nt[ILP];
        float part[ILP];
#pragma unroll
        for (int u = 0; u < ILP; ++u) {
            i0[u] = __shfl_sync(0xffffffffu, li0, (u0 + u) & 31);
            i1[u] = __shfl_sync(0xffffffffu, li1, (u0 + u) & 31);
        }
#pragma unroll
        for (int u = 0; u < ILP; ++u) {
            load_row<MODEL, NCH2>(nh[u], a.ent + i0[u] * d, P, lane);
            load_row<MODEL, NCH2>(nt[u], a.ent + i1[u] * d, P, lane);
        }
#pragma unroll
        for (int u = 0; u < ILP; ++u) part[u] = partial_score<MODEL, NCH2>(nh[u], nt[u], rb);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int u = 0; u < ILP; ++u) part[u] += __shfl_xor_sync(0xffffffffu, part[u], o);
        }
#pragma unroll
        for (int u = 0; u < ILP; ++u) {
            const long long kk = k0 + u;
            if (kk >= we) continue;   // warp-uniform
            const float sc = finish_score<MODEL>(part[u]);
            if (a.neg_scores && lane == 0) a.neg_scores[b * a.k + kk] = sc;
            float w;
            if (a.loss == BLP_LOSS_MARGIN) {
                // models.py:252-253: m = fl(fl(1 - pos) + neg); entries with m < 0 are zeroed (m == 0 keeps its gradient)
                const float m = one_minus_pos + sc;
                const bool keep = !(m < 0.f);
                if (keep) loss_part += m;
                w = keep ? inv_bk : 0.f;
                wsum += w;
            } else {
                // models.py:258: softplus(neg).mean() / 2
                loss_part += softplus_f(sc);
                w = 0.5f * inv_bk * softplus_grad_f(sc);
            }
            if (GRAD && w != 0.f) {
                const float ws = w * half;
                const bool own_h = (i0[u] == 2 * b), own_t = (i1[u] == 2 * b + 1);   // warp-uniform
                if (own_h && own_t) {
                    accumulate<MODEL, NCH2>(ws, nh[u], nt[u], rb, acc.h, acc.t, acc.r);
                } else if (own_h) {
                    Row<NCH2> g;
#pragma unroll
                    for (int c = 0; c < NCH2; ++c) g.lo[c] = g.hi[c] = make_float2(0.f, 0.f);
                    accumulate<MODEL, NCH2>(ws, nh[u], nt[u], rb, acc.h, g, acc.r);
                    flush_row<MODEL, NCH2>(a.grad_ent + i1[u] * d, g, P, lane);
                } else if (own_t) {
                    Row<NCH2> g;
#pragma unroll
                    for (int c = 0; c < NCH2; ++c) g.lo[c] = g.hi[c] = make_float2(0.f, 0.f);
                    accumulate<MODEL, NCH2>(ws, nh[u], nt[u], rb, g, acc.t, acc.r);
                    flush_row<MODEL, NCH2>(a.grad_ent + i0[u] * d, g, P, lane);
                } else {
                    Row<NCH2> g0, g1;
#pragma unroll
                    for (int c = 0; c < NCH2; ++c) g0.lo[c] = g0.hi[c] = g1.lo[c] = g1.hi[c] = make_float2(0.f, 0.f);
                    accumulate<MODEL, NCH2>(ws, nh[u], nt[u], rb, g0, g1, acc.r);
                    flush_row<MODEL, NCH2>(a.grad_ent + i0[u] * d, g0, P, lane);
                    flush_row<MODEL, NCH2>(a.grad_ent + i1[u] * d, g1, P, lane);
                }
            }
        }
    }
    }

    // positive-side gradient: margin d/dpos = -sum_k w_bk (each warp adds its share); nll: -sigmoid(-pos)/(2B)
    const bool lead = (slice == 0 && warp == 0);
    if (GRAD) {
        float wpos = (a.loss == BLP_LOSS_MARGIN) ? -wsum : (lead ? -0.5f * softplus_grad_f(-pos) / (float)a.b : 0.f);
        if (wpos != 0.f) accumulate<MODEL, NCH2>(wpos * half, hb, tb, rb, acc.h, acc.t, acc.r);
        if (lead && a.regularizer > 0.f) {
            // models.py:59-62, 261-266: d/dx of regularizer * mean(x^2) / 3
            const float cr = a.regularizer * 2.0f / (3.0f * (float)a.b * (float)d);
#pragma unroll
            for (int c = 0; c < NCH2; ++c) {
                acc.h.lo[c].x += cr * hb.lo[c].x; acc.h.lo[c].y += cr * hb.lo[c].y;
                acc.t.lo[c].x += cr * tb.lo[c].x; acc.t.lo[c].y += cr * tb.lo[c].y;
                acc.r.lo[c].x += cr * rb.lo[c].x; acc.r.lo[c].y += cr * rb.lo[c].y;
                acc.h.hi[c].x += cr * hb.hi[c].x; acc.h.hi[c].y += cr * hb.hi[c].y;
                acc.t.hi[c].x += cr * tb.hi[c].x; acc.t.hi[c].y += cr * tb.hi[c].y;
                acc.r.hi[c].x += cr * rb.hi[c].x; acc.r.hi[c].y += cr * rb.hi[c].y;
            }
        }
        flush_row<MODEL, NCH2>(a.grad_ent + (2 * b) * d, acc.h, P, lane);
        flush_row<MODEL, NCH2>(a.grad_ent + (2 * b + 1) * d, acc.t, P, lane);
        flush_row<MODEL, NCH2>(a.grad_rel + rel * d, acc.r, P, lane);
    }

    // ---- loss: per-CTA partial, deterministic final reduction by the last CTA
    __shared__ float s_part[kTrainWarps];
    __shared__ bool s_last;
    float mine;
    if (a.loss == BLP_LOSS_MARGIN) mine = loss_part * inv_bk;
    else mine = loss_part * 0.5f * inv_bk;
    if (lead) {
        if (a.loss == BLP_LOSS_NLL) mine += 0.5f * softplus_f(-pos) / (float)a.b;
        if (a.regularizer > 0.f) {
            float sq = 0.f;
#pragma unroll
            for (int c = 0; c < NCH2; ++c) {
                sq += hb.lo[c].x * hb.lo[c].x + hb.lo[c].y * hb.lo[c].y + hb.hi[c].x * hb.hi[c].x + hb.hi[c].y * hb.hi[c].y;
                sq += tb.lo[c].x * tb.lo[c].x + tb.lo[c].y * tb.lo[c].y + tb.hi[c].x * tb.hi[c].x + tb.hi[c].y * tb.hi[c].y;
                sq += rb.lo[c].x * rb.lo[c].x + rb.lo[c].y * rb.lo[c].y + rb.hi[c].x * rb.hi[c].x + rb.hi[c].y * rb.hi[c].y;
            }
            sq = warp_sum(sq);
            mine += a.regularizer * sq / (3.0f * (float)a.b * (float)d);
        }
        if (lane == 0) a.pos_scores[b] = pos;
    }
    if (lane == 0) s_part[warp] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int w = 0; w < kTrainWarps; ++w) tot += s_part[w];
        a.partials[blockIdx.x] = tot;
        __threadfence();
        const unsigned int done = atomicAdd(a.counter, 1u);
        s_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && warp == 0) {
        __threadfence();
        double tot = 0.0;
        for (unsigned int i = lane; i < gridDim.x; i += 32) tot += (double)__ldcg(a.partials + i);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if (lane == 0) {
            *a.loss_out = (float)tot;
            *a.counter = 0u;   // leave the workspace zeroed for the next call
        }
    }
}

// ---- d == 128: lane-group mapping ------------------------------------------------------------------------
// ncu on the generic kernel above (B = 1024, K = 512): 148 warp instructions per negative, IPC 2.4 of 4, one CTA per
// SM, 0.97 M REDG.v2 warp instructions (33 M lane-ops at ~1 lane-op / clk / SM).  With one warp per negative a lane owns
// only 4 of the 128 dims, so the per-negative bookkeeping (index shuffles, 64-bit address arithmetic, loss weight, own /
// foreign row decisions) is replicated 32-fold.  Here a warp is 4 groups of 8 lanes; a group scores one negative (16
// dims per lane: four 16-byte loads per row, 3 shuffle steps per reduction), every bookkeeping instruction serves 4
// negatives, and the gradient scatter is REDG.v4 (half the reduction lane-ops).
#ifndef BLP_TRAIN128_UNROLL
#define BLP_TRAIN128_UNROLL 1
#endif
struct V16 {
    float v[16];
};

// lane gl of a group owns four 16-byte chunks of a 128-float row, interleaved with the other lanes so that every vector
// load / reduction instruction of a group covers 128 contiguous bytes (whole sectors): chunks gl, gl + 8, gl + 16,
// gl + 24 (TransE, DistMult), or chunks gl, gl + 8 of BOTH halves (ComplEx / SimplE: v[0..7] first half, v[8..15] second)
template <int MODEL>
__device__ __forceinline__ int v16_offset(int gl, int c) {      // float offset of the c-th 16-byte chunk
    if (TM<MODEL>::kHalves) return (c >> 1) * 64 + 4 * (gl + 8 * (c & 1));
    return 4 * (gl + 8 * c);
}
template <int MODEL>
__device__ __forceinline__ void load_v16(V16 &x, const float *__restrict__ row, int gl) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float4 q = *reinterpret_cast<const float4 *>(row + v16_offset<MODEL>(gl, c));
        x.v[4 * c] = q.x; x.v[4 * c + 1] = q.y; x.v[4 * c + 2] = q.z; x.v[4 * c + 3] = q.w;
    }
}
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
template <int MODEL>
__device__ __forceinline__ void red_v16(float *__restrict__ row, const V16 &g, int gl, bool pred) {
    if (!pred) return;
#pragma unroll
    for (int c = 0; c < 4; ++c)
        red_add_v4(row + v16_offset<MODEL>(gl, c), g.v[4 * c], g.v[4 * c + 1], g.v[4 * c + 2], g.v[4 * c + 3]);
}
template <int MODEL>
__device__ __forceinline__ float partial16(const V16 &h, const V16 &t, const V16 &r) {
    float s = 0.f;
    constexpr int n = TM<MODEL>::kHalves ? 8 : 16;
#pragma unroll
    for (int i = 0; i < n; ++i)
        s += term1<MODEL>(h.v[i], h.v[(i + 8) & 15], t.v[i], t.v[(i + 8) & 15], r.v[i], r.v[(i + 8) & 15]);
    return s;
}
__device__ __forceinline__ float group_sum(float v) {           // over the 8 lanes of a group
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}
// gh / gt / gr += w * d score / d (h, t, r) on this lane's 16 floats
template <int MODEL>
__device__ __forceinline__ void grad16(float w, const V16 &h, const V16 &t, const V16 &r, V16 &gh, V16 &gt, V16 &gr) {
    constexpr int n = TM<MODEL>::kHalves ? 8 : 16;
#pragma unroll
    for (int i = 0; i < n; ++i) {
        const int j = (i + 8) & 15;
        float a, b2, c2, d2, e, f;
        grad1<MODEL>(w, h.v[i], h.v[j], t.v[i], t.v[j], r.v[i], r.v[j], a, b2, c2, d2, e, f);
        gh.v[i] = a; gt.v[i] = c2; gr.v[i] = e;
        if (TM<MODEL>::kHalves) { gh.v[j] = b2; gt.v[j] = d2; gr.v[j] = f; }
    }
}

template <int MODEL, bool GRAD>
__global__ void __launch_bounds__(kTrainThreads) train128_kernel(const TrainArgs a) {
    constexpr int d = 128;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gid = lane >> 3, gl = lane & 7;
    const long long b = blockIdx.x / a.slices;
    const int slice = blockIdx.x % a.slices;
    const float half = (MODEL == BLP_MODEL_SIMPLE) ? 0.5f : 1.0f;
    const long long nb2 = 2 * a.b;

    long long rel = a.rels[b];
    if (rel < 0 || rel >= a.num_rel) {
        if (threadIdx.x == 0) *a.err_flag = 1;
        rel = 0;
    }
    V16 rb;
    load_v16<MODEL>(rb, a.rel_weight + rel * d, gl);
    float pos;
    {
        V16 hb, tb;
        load_v16<MODEL>(hb, a.ent + (2 * b) * d, gl);
        load_v16<MODEL>(tb, a.ent + (2 * b + 1) * d, gl);
        pos = finish_score<MODEL>(group_sum(partial16<MODEL>(hb, tb, rb)));
    }
    const float one_minus_pos = 1.0f - pos;
    const float inv_bk = 1.0f / (float)(a.b * a.k);
    V16 acc_h, acc_t, acc_r;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc_h.v[i] = acc_t.v[i] = acc_r.v[i] = 0.f;
    float loss_part = 0.f, wsum = 0.f;           // per group (the same value in its 8 lanes)

    const long long kps = (a.k + a.slices - 1) / a.slices;
    const long long kb = slice * kps;
    const long long ke = min(a.k, kb + kps);
    const long long *nrow = a.neg_idx + b * a.s0;
    const long long per_warp = (ke - kb + kTrainWarps - 1) / kTrainWarps;
    const long long wb = kb + (long long)warp * per_warp, we = min(ke, wb + per_warp);
    for (long long base = wb; base < we; base += 32) {
        // the index pairs of 32 negatives with one load per lane, handed out by shuffle (4 negatives per step)
        long long li0 = 2 * b, li1 = 2 * b + 1;
        if (base + lane < we) {
            li0 = nrow[(base + lane) * a.s1];
            li1 = nrow[(base + lane) * a.s1 + a.s2];
            if (li0 < 0 || li0 >= nb2 || li1 < 0 || li1 >= nb2) {
                *a.err_flag = 1;
                li0 = 2 * b; li1 = 2 * b + 1;
            }
        }
        const int cnt = (int)min(32ll, we - base);
        constexpr int kUnroll = BLP_TRAIN128_UNROLL;
#pragma unroll kUnroll
        for (int u0 = 0; u0 < cnt; u0 += 4) {
            const int u = u0 + gid;
            const bool valid = u < cnt;
            const long long i0 = __shfl_sync(0xffffffffu, li0, u & 31), i1 = __shfl_sync(0xffffffffu, li1, u & 31);
            V16 nh, nt;
            load_v16<MODEL>(nh, a.ent + i0 * d, gl);
            load_v16<MODEL>(nt, a.ent + i1 * d, gl);
            const float sc = finish_score<MODEL>(group_sum(partial16<MODEL>(nh, nt, rb)));
            if (a.neg_scores && valid && gl == 0) a.neg_scores[b * a.k + base + u] = sc;
            float w;
            if (a.loss == BLP_LOSS_MARGIN) {
                // models.py:252-253: m = fl(fl(1 - pos) + neg); entries with m < 0 are zeroed (m == 0 keeps its gradient)
                const float m = one_minus_pos + sc;
                const bool keep = valid && !(m < 0.f);
                if (keep) loss_part += m;
                w = keep ? inv_bk : 0.f;
                wsum += w;
            } else {
                // models.py:258: softplus(neg).mean() / 2
                if (valid) loss_part += softplus_f(sc);
                w = valid ? 0.5f * inv_bk * softplus_grad_f(sc) : 0.f;
            }
            if (GRAD && __any_sync(0xffffffffu, w != 0.f)) {
                const bool own_h = (i0 == 2 * b), own_t = (i1 == 2 * b + 1);
                V16 gh, gt, gr;
                grad16<MODEL>(w * half, nh, nt, rb, gh, gt, gr);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    acc_r.v[i] += gr.v[i];
                    acc_h.v[i] += own_h ? gh.v[i] : 0.f;
                    acc_t.v[i] += own_t ? gt.v[i] : 0.f;
                }
                // rows sampled as corrupting entities: vector reductions into grad_ent
                red_v16<MODEL>(a.grad_ent + i0 * d, gh, gl, !own_h && w != 0.f);
                red_v16<MODEL>(a.grad_ent + i1 * d, gt, gl, !own_t && w != 0.f);
            }
        }
    }

    // positive-side gradient: margin d/dpos = -sum_k w_bk (each group adds its share); nll: -sigmoid(-pos)/(2B)
    const bool lead = (slice == 0 && warp == 0);
    V16 hb, tb;
    load_v16<MODEL>(hb, a.ent + (2 * b) * d, gl);
    load_v16<MODEL>(tb, a.ent + (2 * b + 1) * d, gl);
    if (GRAD) {
        const float wpos = (a.loss == BLP_LOSS_MARGIN) ? -wsum : ((lead && gid == 0) ? -0.5f * softplus_grad_f(-pos) / (float)a.b : 0.f);
        {
            V16 gh, gt, gr;
            grad16<MODEL>(wpos * half, hb, tb, rb, gh, gt, gr);
#pragma unroll
            for (int i = 0; i < 16; ++i) { acc_h.v[i] += gh.v[i]; acc_t.v[i] += gt.v[i]; acc_r.v[i] += gr.v[i]; }
        }
        if (lead && gid == 0 && a.regularizer > 0.f) {
            // models.py:59-62, 261-266: d/dx of regularizer * mean(x^2) / 3
            const float cr = a.regularizer * 2.0f / (3.0f * (float)a.b * (float)d);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                acc_h.v[i] += cr * hb.v[i]; acc_t.v[i] += cr * tb.v[i]; acc_r.v[i] += cr * rb.v[i];
            }
        }
        // the 4 groups of the warp hold partial rows for the same 16 floats per lane position: fold them, group 0 flushes
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            acc_h.v[i] += __shfl_xor_sync(0xffffffffu, acc_h.v[i], 8);  acc_h.v[i] += __shfl_xor_sync(0xffffffffu, acc_h.v[i], 16);
            acc_t.v[i] += __shfl_xor_sync(0xffffffffu, acc_t.v[i], 8);  acc_t.v[i] += __shfl_xor_sync(0xffffffffu, acc_t.v[i], 16);
            acc_r.v[i] += __shfl_xor_sync(0xffffffffu, acc_r.v[i], 8);  acc_r.v[i] += __shfl_xor_sync(0xffffffffu, acc_r.v[i], 16);
        }
        red_v16<MODEL>(a.grad_ent + (2 * b) * d, acc_h, gl, gid == 0);
        red_v16<MODEL>(a.grad_ent + (2 * b + 1) * d, acc_t, gl, gid == 0);
        red_v16<MODEL>(a.grad_rel + rel * d, acc_r, gl, gid == 0);
    }

    // ---- loss: per-CTA partial, deterministic final reduction by the last CTA
    __shared__ float s_part[kTrainWarps];
    __shared__ bool s_last;
    float mine = (gl == 0) ? loss_part : 0.f;                 // one lane per group carries the group's partial
    mine += __shfl_xor_sync(0xffffffffu, mine, 8);
    mine += __shfl_xor_sync(0xffffffffu, mine, 16);
    mine *= (a.loss == BLP_LOSS_MARGIN) ? inv_bk : 0.5f * inv_bk;
    if (lead) {
        if (a.loss == BLP_LOSS_NLL) mine += 0.5f * softplus_f(-pos) / (float)a.b;
        if (a.regularizer > 0.f) {
            float sq = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) sq += hb.v[i] * hb.v[i] + tb.v[i] * tb.v[i] + rb.v[i] * rb.v[i];
            sq = group_sum(sq);
            mine += a.regularizer * sq / (3.0f * (float)a.b * (float)d);
        }
        if (lane == 0) a.pos_scores[b] = pos;
    }
    if (lane == 0) s_part[warp] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int w2 = 0; w2 < kTrainWarps; ++w2) tot += s_part[w2];
        a.partials[blockIdx.x] = tot;
        __threadfence();
        const unsigned int done = atomicAdd(a.counter, 1u);
        s_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && warp == 0) {
        __threadfence();
        double tot = 0.0;
        for (unsigned int i = lane; i < gridDim.x; i += 32) tot += (double)__ldcg(a.partials + i);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if (lane == 0) {
            *a.loss_out = (float)tot;
            *a.counter = 0u;   // leave the workspace zeroed for the next call
        }
    }
}

template <int MODEL>
static int launch_train128(const TrainArgs &a, bool grad, cudaStream_t st) {
    const unsigned grid = (unsigned)(a.b * a.slices);
    prof_begin(2, st);
    if (grad) train128_kernel<MODEL, true><<<grid, kTrainThreads, 0, st>>>(a);
    else train128_kernel<MODEL, false><<<grid, kTrainThreads, 0, st>>>(a);
    prof_end(2, st);
    count_launch();
    return check_cuda(cudaGetLastError(), "train128_kernel launch");
}

template <int MODEL, int NCH2, int ILP>
static int launch_train(const TrainArgs &a, bool grad, cudaStream_t st) {
    const unsigned grid = (unsigned)(a.b * a.slices);
    prof_begin(2, st);
    if (grad) train_kernel<MODEL, NCH2, ILP, true><<<grid, kTrainThreads, 0, st>>>(a);
    else train_kernel<MODEL, NCH2, ILP, false><<<grid, kTrainThreads, 0, st>>>(a);
    prof_end(2, st);
    count_launch();
    return check_cuda(cudaGetLastError(), "train_kernel launch");
}

#ifndef BLP_TRAIN128
#define BLP_TRAIN128 1
#endif
template <int MODEL>
static int dispatch_train(const TrainArgs &a, bool grad, cudaStream_t st) {
    // d = 128 (every BLP script) with many negatives per row: lane-group kernel (rows must be 16-byte aligned for its
    // vector loads / reductions).  Measured (TransE, margin): B = 1024, K = 512: 106 vs 121 us; B = 64, K = 512: 16.4 vs
    // 17.0 us; with K = 64 a warp has only 4 negatives and the per-warp epilogue of this kernel costs more than it
    // saves (B = 8192, K = 64: 338 vs 270 us), so those shapes stay on the warp-per-negative kernel
    if (BLP_TRAIN128 && a.d == 128 && a.k >= 256 && ((reinterpret_cast<uintptr_t>(a.ent) | reinterpret_cast<uintptr_t>(a.rel_weight) |
                                        reinterpret_cast<uintptr_t>(a.grad_ent) | reinterpret_cast<uintptr_t>(a.grad_rel)) & 15u) == 0)
        return launch_train128<MODEL>(a, grad, st);
    const int P = TM<MODEL>::kHalves ? a.d / 2 : a.d;
    const int need = (P + 63) / 64;
#ifndef BLP_TRAIN_ILP
#define BLP_TRAIN_ILP 4
#endif
    if (need <= 1) return launch_train<MODEL, 1, BLP_TRAIN_ILP>(a, grad, st);
    if (need <= 2) return launch_train<MODEL, 2, BLP_TRAIN_ILP>(a, grad, st);
    if (need <= 4) return launch_train<MODEL, 4, 2>(a, grad, st);
    if (need <= 6) return launch_train<MODEL, 6, 2>(a, grad, st);
    if (need <= 12) return launch_train<MODEL, 12, 1>(a, grad, st);
    if (need <= 16) return launch_train<MODEL, 16, 1>(a, grad, st);
    set_error("fused compute_loss supports d <= %d for this model (got %d)", TM<MODEL>::kHalves ? 2048 : 1024, a.d);
    return BLP_EDIM;
}

}  // namespace blp

using namespace blp;

extern "C" int64_t blp_train_workspace_bytes(int64_t b, int64_t k) {
    (void)k;
    if (b < 0) return 0;
    return 64 + 4 * b * 8;   // counter + error flag + one float partial per CTA (<= 8 slices per row)
}

extern "C" int blp_train_loss(int model, int loss, const float *ent_embs, const float *rel_weight, const int64_t *rels,
                              int64_t num_rel, const int64_t *neg_idx, int64_t s0, int64_t s1, int64_t s2, int64_t b,
                              int64_t k, int d, float regularizer, float *loss_out, float *pos_scores,
                              float *neg_scores, float *grad_ent, float *grad_rel_weight, void *workspace, void *stream) {
    reset_launch_count();
    if (model < 0 || model > 3) { set_error("unknown relational model id %d", model); return BLP_EINVAL; }
    if (loss != BLP_LOSS_MARGIN && loss != BLP_LOSS_NLL) { set_error("unknown loss id %d", loss); return BLP_EINVAL; }
    if (b <= 0 || k <= 0 || num_rel <= 0) { set_error("b, k and num_rel must be positive"); return BLP_EINVAL; }
    if (d <= 0 || (d & 1)) { set_error("fused compute_loss needs an even d (got %d)", d); return BLP_EDIM; }
    if ((model == BLP_MODEL_COMPLEX || model == BLP_MODEL_SIMPLE) && (d & 3)) {
        set_error("fused compute_loss needs d %% 4 == 0 for complex/simple (got %d)", d);
        return BLP_EDIM;
    }
    if (!ent_embs || !rel_weight || !rels || !neg_idx || !loss_out || !pos_scores || !workspace) {
        set_error("null pointer argument");
        return BLP_EINVAL;
    }
    if ((grad_ent == nullptr) != (grad_rel_weight == nullptr)) { set_error("grad_ent and grad_rel_weight must both be given or both NULL"); return BLP_EINVAL; }
    if ((reinterpret_cast<uintptr_t>(ent_embs) | reinterpret_cast<uintptr_t>(rel_weight)) & 7u) {
        set_error("ent_embs / rel_weight must be 8-byte aligned");
        return BLP_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const bool grad = grad_ent != nullptr;
    if (grad) {
        const size_t n_ent = (size_t)(2 * b) * d, n_rel = (size_t)num_rel * d;
        if (grad_rel_weight == grad_ent + n_ent) {   // one allocation (the Python wrapper's layout): one memset
            BLP_CUDA(cudaMemsetAsync(grad_ent, 0, sizeof(float) * (n_ent + n_rel), st));
        } else {
            BLP_CUDA(cudaMemsetAsync(grad_ent, 0, sizeof(float) * n_ent, st));
            BLP_CUDA(cudaMemsetAsync(grad_rel_weight, 0, sizeof(float) * n_rel, st));
        }
    }
    TrainArgs a{};
    a.ent = ent_embs; a.rel_weight = rel_weight; a.rels = (const long long *)rels; a.neg_idx = (const long long *)neg_idx;
    a.s0 = s0; a.s1 = s1; a.s2 = s2; a.b = b; a.k = k; a.num_rel = num_rel; a.d = d; a.loss = loss; a.regularizer = regularizer;
    a.loss_out = loss_out; a.pos_scores = pos_scores; a.neg_scores = neg_scores; a.grad_ent = grad_ent; a.grad_rel = grad_rel_weight;
    a.counter = reinterpret_cast<unsigned int *>(workspace);
    a.err_flag = reinterpret_cast<int *>(workspace) + 1;
    a.partials = reinterpret_cast<float *>(workspace) + 16;
    // enough CTAs to cover the SMs about twice, at least one 16-warp pass of negatives per CTA
    // one wave of CTAs: every extra CTA repeats the positive-row prologue and the register-accumulator flush, and a
    // second wave costs a full latency chain (B=64, K=512: 2 slices = 20.5 us, 8 slices = 30 us)
    int slices = 1;
    while (slices < 8 && b * slices * 2 <= 160 && k / (slices * 2) >= 64) slices *= 2;
    {
        static const int forced = []() {                  // tuning aid, read once (not per step)
            const char *e = getenv("BLP_TRAIN_SLICES");
            return (e && atoi(e) >= 1 && atoi(e) <= 8) ? atoi(e) : 0;
        }();
        if (forced) slices = forced;
    }
    a.slices = slices;
    switch (model) {
    case BLP_MODEL_TRANSE: return dispatch_train<BLP_MODEL_TRANSE>(a, grad, st);
    case BLP_MODEL_DISTMULT: return dispatch_train<BLP_MODEL_DISTMULT>(a, grad, st);
    case BLP_MODEL_COMPLEX: return dispatch_train<BLP_MODEL_COMPLEX>(a, grad, st);
    default: return dispatch_train<BLP_MODEL_SIMPLE>(a, grad, st);
    }
}
