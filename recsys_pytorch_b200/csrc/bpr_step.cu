// Fused BPR-MF training step for sm_100a.
//
// Replaces, as ONE kernel per batch, the reference's
//   data/generators.py:168-201   negative sampling (uniform over non-positives)
//   models/MF.py:32-42,99-107    4 embedding gathers, 2 mul+sum, sub/sigmoid/log/mean
//   models/MF.py:67-68           autograd backward (dense [U,d],[I,d] grads) + optimizer.step
// with: [counter-RNG sample against the CSR row] -> gather u/i/j rows once ->
// warp-reduced x = u.(vi-vj) -> g = -sigmoid(-x)/B -> vector-atomic scatter of
// the three row updates.  The dense gradient is never materialised.
//
// Work decomposition: a warp owns a CHUNK of <=32 consecutive triples whose ids
// live one-per-lane (coalesced id loads / lane-parallel sampling); rows are then
// processed by sub-groups of G lanes (G*16 B >= row bytes when d <= 128), i.e.
// 32/G triples at a time, ids broadcast with shuffles.
//
// Two gather paths:
//   LDG  : each lane loads its float4 of the three rows (prefetching the next
//          triple while the current one is reduced).
//   TMA  : warp-private ring of shared-memory stages filled by cp.async.bulk
//          (UBLKCP) row copies that complete on a warp-private mbarrier - no CTA
//          level synchronisation at all; many rows in flight per warp.
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "sampler.cuh"

namespace b200 {

struct BprParams {
    b200rec_bpr_args a;
    float invB;
    int chunk;       // triples per warp-chunk (multiple of 32/G, <= 32)
    int stages;      // TMA path: shared-memory stages per warp
    int64_t n_chunks;
    unsigned int *work;   // fast path: chunk counter (zeroed before the launch) for dynamic distribution, or NULL
};

// dL/dx of -log(sigmoid(x))/B the way the reference's fp32 autograd evaluates it (models/MF.py:105):
// s = fl(1/(1+fl(exp(-x)))), grad = -(1 - s)/B.  It saturates to exactly 0 for x > ~16.6 (s rounds to 1),
// which matters under Adam (a 1e-10 gradient would still move a weight by ~lr).  For x < -88 the
// reference produces NaN; here s -> 0 and the gradient is the finite limit -1/B.
__device__ __forceinline__ float bpr_grad(float x, float invB) {
    const float s = 1.f / (1.f + expf(-x));
    return -(1.f - s) * invB;
}

__device__ __forceinline__ float softplus_neg(float x) {  // -log(sigmoid(x)) = log(1+exp(-x)), stable
    return x > 0.f ? log1pf(expf(-x)) : (-x + log1pf(expf(x)));
}

// four floats -> four bf16 (round to nearest), added atomically as two bf16x2 words (REDG.E.ADD.BF16x4)
__device__ __forceinline__ void red4_bf16(__nv_bfloat16 *p, float4 v) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    asm volatile("red.relaxed.gpu.global.add.noftz.v2.bf16x2 [%0], {%1,%2};" ::"l"(p),
                 "r"(*reinterpret_cast<const uint32_t *>(&lo)), "r"(*reinterpret_cast<const uint32_t *>(&hi))
                 : "memory");
}

template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// scatter of one float4 column-chunk of the three rows of a triple
template <int SINK>
__device__ __forceinline__ void sink_chunk(const b200rec_bpr_args &a, int64_t t, int tu, int ti, int tj, int q,
                                           float g, float regB, float4 ru, float4 ri, float4 rj, bool uniq) {
    const float sc = (SINK == B200REC_SINK_GRAD) ? 1.f : -a.lr;
    float4 du, di, dj;
    du.x = sc * fmaf(g, ri.x - rj.x, regB * ru.x); du.y = sc * fmaf(g, ri.y - rj.y, regB * ru.y);
    du.z = sc * fmaf(g, ri.z - rj.z, regB * ru.z); du.w = sc * fmaf(g, ri.w - rj.w, regB * ru.w);
    di.x = sc * fmaf(g, ru.x, regB * ri.x); di.y = sc * fmaf(g, ru.y, regB * ri.y);
    di.z = sc * fmaf(g, ru.z, regB * ri.z); di.w = sc * fmaf(g, ru.w, regB * ri.w);
    dj.x = sc * fmaf(-g, ru.x, regB * rj.x); dj.y = sc * fmaf(-g, ru.y, regB * rj.y);
    dj.z = sc * fmaf(-g, ru.z, regB * rj.z); dj.w = sc * fmaf(-g, ru.w, regB * rj.w);
    const int64_t ld = a.ld;
    if (SINK == B200REC_SINK_UPDATE) {
        if (a.udelta) {                                   // item-sharded: user delta row -> exchange buffer [B, ld]
            st4(a.udelta + t * ld + q * 4, du);
        } else {
            float *pu = a.U + (int64_t)tu * ld + q * 4;
            if (uniq) st4(pu, make_float4(ru.x + du.x, ru.y + du.y, ru.z + du.z, ru.w + du.w));
            else red4(pu, du);
        }
        // user-sharded: item deltas accumulate in a dense exchange buffer instead of the replica
        if (a.flags & B200REC_F_ITEM_DELTA_BF16) {   // ... held in bf16: half the L2 footprint and half the wire bytes
            __nv_bfloat16 *vb = reinterpret_cast<__nv_bfloat16 *>(a.gV);
            red4_bf16(vb + (int64_t)ti * ld + q * 4, di);
            red4_bf16(vb + (int64_t)tj * ld + q * 4, dj);
        } else {
            float *vdst = (a.flags & B200REC_F_ITEM_DELTA) ? a.gV : a.V;
            red4(vdst + (int64_t)(ti - a.item_lo) * ld + q * 4, di);
            red4(vdst + (int64_t)(tj - a.item_lo) * ld + q * 4, dj);
        }
    } else if (SINK == B200REC_SINK_STAGE) {
        float *s = a.stage + (t * 3) * ld + q * 4;
        st4(s, du); st4(s + ld, di); st4(s + 2 * ld, dj);
    } else if (SINK == B200REC_SINK_GRAD) {
        red4(a.gU + (int64_t)tu * ld + q * 4, du);
        red4(a.gV + (int64_t)ti * ld + q * 4, di);
        red4(a.gV + (int64_t)tj * ld + q * 4, dj);
    }
}

// ---------------------------------------------------------------------------
// LDG gather path
// ---------------------------------------------------------------------------
template <int G, int CPL, int SINK>
__global__ void __launch_bounds__(256) bpr_step_ldg_kernel(const BprParams p) {
    const b200rec_bpr_args &a = p.a;
    constexpr int TPW = 32 / G;
    const int lane = threadIdx.x & 31;
    const int sl = lane % G, sg = lane / G;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int d4 = a.ld >> 2;
    const int64_t ld = a.ld;
    const float regB = a.reg * p.invB;
    const bool uniq = (a.flags & B200REC_F_USERS_UNIQUE) != 0;
    const int iters = p.chunk / TPW;
    float loss_local = 0.f;

    for (int64_t c = warp_global; c < p.n_chunks; c += n_warps) {
        const int64_t t_lane = c * p.chunk + lane;
        bool valid = (lane < p.chunk) && (t_lane < a.B);
        int u, i, j;
        fetch_triple(a, t_lane, valid, u, i, j);

        float4 nu[CPL], ni[CPL], nj[CPL];
        int tu, ti, tj, tv;
        auto issue = [&](int it) {
            const int src = it * TPW + sg;
            tu = __shfl_sync(0xffffffffu, u, src);
            ti = __shfl_sync(0xffffffffu, i, src);
            tj = __shfl_sync(0xffffffffu, j, src);
            tv = __shfl_sync(0xffffffffu, (int)valid, src);
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                const int q = sl + k * G;
                if (tv && q < d4) {
                    nu[k] = ld4(a.U + (int64_t)tu * ld + q * 4);
                    ni[k] = ld4(a.V + (int64_t)(ti - a.item_lo) * ld + q * 4);
                    nj[k] = ld4(a.V + (int64_t)(tj - a.item_lo) * ld + q * 4);
                } else {
                    nu[k] = ni[k] = nj[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        };
        issue(0);
        for (int it = 0; it < iters; ++it) {
            float4 ru[CPL], ri[CPL], rj[CPL];
            const int cu = tu, ci = ti, cj = tj, cv = tv;
#pragma unroll
            for (int k = 0; k < CPL; ++k) { ru[k] = nu[k]; ri[k] = ni[k]; rj[k] = nj[k]; }
            if (it + 1 < iters) issue(it + 1);  // next triple's rows in flight while this one reduces
            float part = 0.f;
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                part = fmaf(ru[k].x, ri[k].x - rj[k].x, part);
                part = fmaf(ru[k].y, ri[k].y - rj[k].y, part);
                part = fmaf(ru[k].z, ri[k].z - rj[k].z, part);
                part = fmaf(ru[k].w, ri[k].w - rj[k].w, part);
            }
            const float x = group_sum<G>(part);
            const float g = bpr_grad(x, p.invB);
            const int64_t t = c * p.chunk + it * TPW + sg;
            if (cv) {
                if (sl == 0) {
                    if (a.loss_sum) loss_local += softplus_neg(x);
                    if (a.x_out) a.x_out[t] = x;
                }
#pragma unroll
                for (int k = 0; k < CPL; ++k) {
                    const int q = sl + k * G;
                    if (q < d4) sink_chunk<SINK>(a, t, cu, ci, cj, q, g, regB, ru[k], ri[k], rj[k], uniq);
                }
            }
        }
    }
    if (a.loss_sum) {
        loss_local = group_sum<32>(loss_local);
        if (lane == 0 && loss_local != 0.f) atomicAdd(a.loss_sum, (double)loss_local);
    }
}

// ---------------------------------------------------------------------------
// TMA (cp.async.bulk) gather path: warp-private stage ring + mbarriers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

constexpr int kTmaWarps = 8;   // warps per CTA
constexpr int kTmaMaxStages = 8;  // stages per warp (fewer when rows are wide)

template <int G, int CPL, int SINK>
__global__ void __launch_bounds__(kTmaWarps * 32) bpr_step_tma_kernel(const BprParams p) {
    const b200rec_bpr_args &a = p.a;
    constexpr int TPW = 32 / G;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int sl = lane % G, sg = lane / G;
    const int64_t ld = a.ld;
    const int d4 = a.ld >> 2;
    const uint32_t row_bytes = (uint32_t)a.ld * 4u;
    const uint32_t stage_bytes = row_bytes * 3u * TPW;
    // layout: [warps][stages] mbarriers (8 B each) then [warps][stages][TPW][3][ld] floats
    const int kTmaStages = p.stages;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw) + wid * kTmaMaxStages;
    const uint32_t data_off = (uint32_t)(kTmaWarps * kTmaMaxStages * 8);
    float *wdata = reinterpret_cast<float *>(smem_raw + data_off + (size_t)wid * kTmaStages * stage_bytes);
    if (lane < kTmaStages) mbar_init(smem_u32(bars + lane), TPW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();

    const int64_t warp_global = (int64_t)blockIdx.x * kTmaWarps + wid;
    const int64_t n_warps = (int64_t)gridDim.x * kTmaWarps;
    const float regB = a.reg * p.invB;
    const bool uniq = (a.flags & B200REC_F_USERS_UNIQUE) != 0;
    const int iters = p.chunk / TPW;
    float loss_local = 0.f;
    uint32_t n_consumed = 0;  // running stage counter -> slot and phase parity

    for (int64_t c = warp_global; c < p.n_chunks; c += n_warps) {
        const int64_t t_lane = c * p.chunk + lane;
        bool valid = (lane < p.chunk) && (t_lane < a.B);
        int u, i, j;
        fetch_triple(a, t_lane, valid, u, i, j);

        // the sub-group leader (sl==0) of triple-slot sg issues that triple's three row copies
        auto issue = [&](int it, uint32_t n_slot) {
            const int src = it * TPW + sg;
            const int tu = __shfl_sync(0xffffffffu, u, src);
            const int ti = __shfl_sync(0xffffffffu, i, src);
            int tj = __shfl_sync(0xffffffffu, j, src);
            int ti_ = ti;
            if (!__shfl_sync(0xffffffffu, (int)valid, src)) { ti_ = a.item_lo; tj = a.item_lo; }  // keep addresses in range
            if (sl == 0) {
                const uint32_t slot = n_slot % kTmaStages;
                const uint32_t bar = smem_u32(bars + slot);
                const uint32_t dst = smem_u32(wdata) + slot * stage_bytes + (uint32_t)sg * 3u * row_bytes;
                mbar_expect_tx(bar, 3u * row_bytes);
                bulk_g2s(dst, a.U + (int64_t)tu * ld, row_bytes, bar);
                bulk_g2s(dst + row_bytes, a.V + (int64_t)(ti_ - a.item_lo) * ld, row_bytes, bar);
                bulk_g2s(dst + 2u * row_bytes, a.V + (int64_t)(tj - a.item_lo) * ld, row_bytes, bar);
            }
        };
        const int pre = iters < kTmaStages ? iters : kTmaStages;
        for (int it = 0; it < pre; ++it) issue(it, n_consumed + it);

        for (int it = 0; it < iters; ++it) {
            const uint32_t slot = n_consumed % kTmaStages;
            const uint32_t parity = (n_consumed / kTmaStages) & 1u;
            const int src = it * TPW + sg;
            const int cu = __shfl_sync(0xffffffffu, u, src);
            const int ci = __shfl_sync(0xffffffffu, i, src);
            const int cj = __shfl_sync(0xffffffffu, j, src);
            const int cv = __shfl_sync(0xffffffffu, (int)valid, src);
            mbar_wait(smem_u32(bars + slot), parity);
            const float *st = wdata + (size_t)slot * (stage_bytes / 4) + (size_t)sg * 3 * ld;
            float4 ru[CPL], ri[CPL], rj[CPL];
            float part = 0.f;
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                const int q = sl + k * G;
                if (q < d4) {
                    ru[k] = ld4(st + q * 4); ri[k] = ld4(st + ld + q * 4); rj[k] = ld4(st + 2 * ld + q * 4);
                } else {
                    ru[k] = ri[k] = rj[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                part = fmaf(ru[k].x, ri[k].x - rj[k].x, part);
                part = fmaf(ru[k].y, ri[k].y - rj[k].y, part);
                part = fmaf(ru[k].z, ri[k].z - rj[k].z, part);
                part = fmaf(ru[k].w, ri[k].w - rj[k].w, part);
            }
            const float x = group_sum<G>(part);  // every lane's smem reads are complete after this exchange
            ++n_consumed;
            __syncwarp();
            if (it + kTmaStages < iters) issue(it + kTmaStages, n_consumed - 1 + kTmaStages);  // refill this slot
            const float g = bpr_grad(x, p.invB);
            const int64_t t = c * p.chunk + it * TPW + sg;
            if (cv) {
                if (sl == 0) {
                    if (a.loss_sum) loss_local += softplus_neg(x);
                    if (a.x_out) a.x_out[t] = x;
                }
#pragma unroll
                for (int k = 0; k < CPL; ++k) {
                    const int q = sl + k * G;
                    if (q < d4) sink_chunk<SINK>(a, t, cu, ci, cj, q, g, regB, ru[k], ri[k], rj[k], uniq);
                }
            }
        }
    }
    if (a.loss_sum) {
        loss_local = group_sum<32>(loss_local);
        if (lane == 0 && loss_local != 0.f) atomicAdd(a.loss_sum, (double)loss_local);
    }
}

// ---------------------------------------------------------------------------
// Lean fast path: SINK_UPDATE, full-warp rows (ld == 128*CPL floats), single device, on-device or
// given triples, optional loss.  Everything loop-invariant lives in registers (the generic kernel
// re-reads ~40 parameter words and re-tests every feature flag per triple: 264 issued instructions per
// triple, 60% issue-slot utilisation in ncu run 3); rows are prefetched two triples ahead; fast
// ex2/rcp/lg2 for the sigmoid and the loss.  ~75 instructions per triple.
// ---------------------------------------------------------------------------
template <int CPL>
struct RowSet {
    float4 u[CPL], i[CPL], j[CPL];
    int tu, ti, tj;
};

#ifndef B200REC_FAST_MINB
#define B200REC_FAST_MINB 3
#endif
// Profiling-only build (tools/build_ablate.sh, -DB200REC_ABLATE): B200REC_ABL=<bits> switches parts of the fast
// kernel off to measure what each costs (results are then WRONG on purpose).  Never compiled into libb200rec.so.
//   1 item REDs become plain stores   2 no item writes   4 no user write   8 no neg row (load+write)   16 no pos row
#ifdef B200REC_ABLATE
__constant__ int c_abl;
#define ABL(bit) (c_abl & (bit))
#else
#define ABL(bit) 0
#endif
// IDELTA: the two item-row updates go to the dense fp32 item-delta buffer gV (user-sharded multi-GPU layout: one
// all-reduce of gV per step) instead of V itself.
template <int CPL, bool UNIQ, bool LOSS, bool IDELTA = false>
__global__ void __launch_bounds__(256, CPL == 1 ? B200REC_FAST_MINB : 2) bpr_step_fast_kernel(const BprParams p) {
    float *__restrict__ const U = p.a.U;
    float *__restrict__ const V = p.a.V;
    float *__restrict__ const VD = IDELTA ? p.a.gV : p.a.V;
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    constexpr int64_t LD = 128 * CPL;
    const float c_g = p.a.lr * p.invB;             // delta = c_g*(1-s) * other  + c_r * self
    const float c_r = -p.a.lr * p.a.reg * p.invB;
    const int chunk = p.chunk, B = p.a.B;
    const int64_t n_chunks = p.n_chunks;
    b200rec_bpr_args a = p.a;                       // private copy: lets the compiler keep fields in registers
    a.out_pos = p.a.out_pos; a.out_neg = p.a.out_neg;
    float loss_local = 0.f;
    // L2 policy (B200REC_F_L2_HINTS): user rows are touched once per step -> evict first; item rows (and the item
    // delta buffer) are the reused working set -> evict last
    const bool hints = (p.a.flags & B200REC_F_L2_HINTS) != 0;
    const uint64_t pol_u = l2_policy_evict_first(), pol_v = l2_policy_evict_last();

    // Chunks are handed out dynamically when p.work is set: with a static split a CTA that starts late (another
    // kernel - NCCL's all-reduce in the multi-GPU step - holds its SM for a while) still owns 1/grid of the batch and
    // the launch lasts (delay + full kernel time); with the counter late CTAs simply take fewer chunks.
    unsigned int *const work = p.work;
    int64_t c = warp_global;
    int64_t c_next = 0;
    if (work) {
        if (lane == 0) c = (int64_t)atomicAdd(work, 1u);
        c = __shfl_sync(0xffffffffu, c, 0);
    }
    for (; c < n_chunks; c = work ? c_next : c + n_warps) {
        if (work) {   // fetch the next chunk index now: its latency hides behind this chunk's rows
            if (lane == 0) c_next = (int64_t)atomicAdd(work, 1u);
        }
        const int64_t t_lane = c * chunk + lane;
        bool valid = (lane < chunk) && (t_lane < B);
        int u, i, j;
        fetch_triple(a, t_lane, valid, u, i, j);
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);

        RowSet<CPL> r0, r1, r2;
        auto load = [&](RowSet<CPL> &r, int it) {
            if (it < chunk) {
                r.tu = __shfl_sync(0xffffffffu, u, it);
                r.ti = __shfl_sync(0xffffffffu, i, it);
                r.tj = __shfl_sync(0xffffffffu, j, it);
                if ((vmask >> it) & 1u) {
                    const float *pu = U + (int64_t)r.tu * LD + lane * 4;
                    const float *pi = V + (int64_t)r.ti * LD + lane * 4;
                    const float *pj = V + (int64_t)r.tj * LD + lane * 4;
#ifdef B200REC_ABLATE
                    if (ABL(8 | 16)) {
#pragma unroll
                        for (int k = 0; k < CPL; ++k) {
                            r.u[k] = ld4(pu + 128 * k);
                            r.i[k] = ABL(16) ? make_float4(0.1f, 0.1f, 0.1f, 0.1f) : ld4(pi + 128 * k);
                            r.j[k] = ABL(8) ? make_float4(0.f, 0.f, 0.f, 0.f) : ld4(pj + 128 * k);
                        }
                    } else
#endif
                    if (hints) {
#pragma unroll
                        for (int k = 0; k < CPL; ++k) {
                            r.u[k] = ld4_hint(pu + 128 * k, pol_u); r.i[k] = ld4_hint(pi + 128 * k, pol_v);
                            r.j[k] = ld4_hint(pj + 128 * k, pol_v);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < CPL; ++k) { r.u[k] = ld4(pu + 128 * k); r.i[k] = ld4(pi + 128 * k); r.j[k] = ld4(pj + 128 * k); }
                    }
                }
            }
        };
        auto compute = [&](const RowSet<CPL> &r, int it) {
            if (it < chunk && ((vmask >> it) & 1u)) {
                float4 df[CPL];
                float part = 0.f;
#pragma unroll
                for (int k = 0; k < CPL; ++k) {
                    df[k] = make_float4(r.i[k].x - r.j[k].x, r.i[k].y - r.j[k].y, r.i[k].z - r.j[k].z, r.i[k].w - r.j[k].w);
                    part = fmaf(r.u[k].x, df[k].x, part); part = fmaf(r.u[k].y, df[k].y, part);
                    part = fmaf(r.u[k].z, df[k].z, part); part = fmaf(r.u[k].w, df[k].w, part);
                }
                const float x = group_sum<32>(part);
                const float s = __frcp_rn(1.f + __expf(-x));       // sigmoid(x); 1 - s saturates like the reference's fp32
                const float a1 = c_g * (1.f - s);                   // = -lr * g
                if (LOSS) loss_local += (x < -15.f) ? -x : -__logf(s);
                float *pu = U + (int64_t)r.tu * LD + lane * 4;
                float *pi = VD + (int64_t)r.ti * LD + lane * 4;
                float *pj = VD + (int64_t)r.tj * LD + lane * 4;
#pragma unroll
                for (int k = 0; k < CPL; ++k) {
                    float4 du, di, dj;
                    du.x = fmaf(a1, df[k].x, c_r * r.u[k].x); du.y = fmaf(a1, df[k].y, c_r * r.u[k].y);
                    du.z = fmaf(a1, df[k].z, c_r * r.u[k].z); du.w = fmaf(a1, df[k].w, c_r * r.u[k].w);
                    di.x = fmaf(a1, r.u[k].x, c_r * r.i[k].x); di.y = fmaf(a1, r.u[k].y, c_r * r.i[k].y);
                    di.z = fmaf(a1, r.u[k].z, c_r * r.i[k].z); di.w = fmaf(a1, r.u[k].w, c_r * r.i[k].w);
                    dj.x = fmaf(-a1, r.u[k].x, c_r * r.j[k].x); dj.y = fmaf(-a1, r.u[k].y, c_r * r.j[k].y);
                    dj.z = fmaf(-a1, r.u[k].z, c_r * r.j[k].z); dj.w = fmaf(-a1, r.u[k].w, c_r * r.j[k].w);
                    if (hints) {
                        if (UNIQ) st4_hint(pu + 128 * k, make_float4(r.u[k].x + du.x, r.u[k].y + du.y, r.u[k].z + du.z, r.u[k].w + du.w), pol_u);
                        else red4_hint(pu + 128 * k, du, pol_u);
                        red4_hint(pi + 128 * k, di, pol_v);
                        red4_hint(pj + 128 * k, dj, pol_v);
                    }
#ifdef B200REC_ABLATE
                    else if (c_abl) {
                        if (!ABL(4)) {
                            if (UNIQ) st4(pu + 128 * k, make_float4(r.u[k].x + du.x, r.u[k].y + du.y, r.u[k].z + du.z, r.u[k].w + du.w));
                            else red4(pu + 128 * k, du);
                        }
                        if (!ABL(2)) {
                            if (ABL(1)) { if (!ABL(16)) st4(pi + 128 * k, di); if (!ABL(8)) st4(pj + 128 * k, dj); }
                            else { if (!ABL(16)) red4(pi + 128 * k, di); if (!ABL(8)) red4(pj + 128 * k, dj); }
                        }
                    }
#endif
                    else {
                        if (UNIQ) st4(pu + 128 * k, make_float4(r.u[k].x + du.x, r.u[k].y + du.y, r.u[k].z + du.z, r.u[k].w + du.w));
                        else red4(pu + 128 * k, du);
                        red4(pi + 128 * k, di);
                        red4(pj + 128 * k, dj);
                    }
                }
            }
        };
        load(r0, 0);
        load(r1, 1);
        for (int it = 0; it < chunk; it += 3) {
            load(r2, it + 2);
            compute(r0, it);
            load(r0, it + 3);
            compute(r1, it + 1);
            load(r1, it + 4);
            compute(r2, it + 2);
        }
        if (work) c_next = __shfl_sync(0xffffffffu, c_next, 0);
    }
    if (LOSS) {
        // every lane accumulated the same per-triple value
        if (lane == 0 && loss_local != 0.f) atomicAdd(p.a.loss_sum, (double)loss_local);
    }
}

// ---------------------------------------------------------------------------
// Group fast path: a row is held by a sub-warp group of G lanes (32/G float4 per lane and row), so a warp works on
// 32/G triples at once.  Why (profiles/r02_bpr_ablation.md): a random 512-byte row gather alone reaches 6.6 TB/s on
// this part, yet the warp-per-row kernel above needs 0.27 ms just to READ the 1M user rows of a step (1.9 TB/s) - it
// is bound by its own dependent chain (one triple per warp at a time: 5 shuffle stages + exp + rcp between the loads
// of consecutive triples), not by memory.  With G = 8 every lane issues 12 independent 16-byte loads per triple
// group, the reduction is 3 shuffle stages, and the scalar part (sigmoid, loss) is paid once per 4 triples.
// PF = register sets loaded ahead of the one being computed.
// ---------------------------------------------------------------------------
template <int CPL>
struct GroupSet {
    float4 u[CPL], i[CPL], j[CPL];
};

template <int G, int F4, int PF, bool UNIQ, bool LOSS, bool IDELTA>
__global__ void __launch_bounds__(256) bpr_step_group_kernel(const BprParams p) {
    constexpr int CPL = F4 / G, TPW = 32 / G, ITERS = 32 / TPW;
    constexpr int64_t LD = F4 * 4;
    float *__restrict__ const U = p.a.U;
    float *__restrict__ const V = p.a.V;
    float *__restrict__ const VD = IDELTA ? p.a.gV : p.a.V;
    const int lane = threadIdx.x & 31, sl = lane % G, sg = lane / G;
    const float c_g = p.a.lr * p.invB;             // delta = c_g*(1-s) * other  + c_r * self
    const float c_r = -p.a.lr * p.a.reg * p.invB;
    const int B = p.a.B;
    const int64_t n_chunks = p.n_chunks;
    b200rec_bpr_args a = p.a;
    float loss_local = 0.f;
    unsigned int *const work = p.work;
    int64_t c = 0, c_next = 0;
    if (lane == 0) c = (int64_t)atomicAdd(work, 1u);
    c = __shfl_sync(0xffffffffu, c, 0);
    for (; c < n_chunks; c = c_next) {
        if (lane == 0) c_next = (int64_t)atomicAdd(work, 1u);
        const int64_t t_lane = c * 32 + lane;
        bool valid = t_lane < B;
        int u, i, j;
        fetch_triple(a, t_lane, valid, u, i, j);
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);

        GroupSet<CPL> set[PF + 1];
        auto load = [&](GroupSet<CPL> &r, int it) {
            const int src = it * TPW + sg;
            const int tu = __shfl_sync(0xffffffffu, u, src);
            const int ti = __shfl_sync(0xffffffffu, i, src);
            const int tj = __shfl_sync(0xffffffffu, j, src);
            if ((vmask >> src) & 1u) {
                const float *pu = U + (int64_t)tu * LD + sl * 4;
                const float *pi = V + (int64_t)ti * LD + sl * 4;
                const float *pj = V + (int64_t)tj * LD + sl * 4;
#pragma unroll
                for (int k = 0; k < CPL; ++k) r.u[k] = ld4(pu + k * G * 4);
#pragma unroll
                for (int k = 0; k < CPL; ++k) r.i[k] = ld4(pi + k * G * 4);
#pragma unroll
                for (int k = 0; k < CPL; ++k) r.j[k] = ld4(pj + k * G * 4);
            }
        };
        auto compute = [&](GroupSet<CPL> &r, int it) {
            const int src = it * TPW + sg;
            const int tu = __shfl_sync(0xffffffffu, u, src);
            const int ti = __shfl_sync(0xffffffffu, i, src);
            const int tj = __shfl_sync(0xffffffffu, j, src);
            const bool ok = (vmask >> src) & 1u;
            float part = 0.f;
            if (ok) {
#pragma unroll
                for (int k = 0; k < CPL; ++k) {
                    // r.i becomes the difference row (v_i - v_j); r.j keeps v_j (v_i = diff + v_j is not needed: the
                    // L2 term of the positive row uses the loaded value kept in `vi` below)
                    part = fmaf(r.u[k].x, r.i[k].x - r.j[k].x, part); part = fmaf(r.u[k].y, r.i[k].y - r.j[k].y, part);
                    part = fmaf(r.u[k].z, r.i[k].z - r.j[k].z, part); part = fmaf(r.u[k].w, r.i[k].w - r.j[k].w, part);
                }
            }
            const float x = group_sum<G>(part);
            if (ok) {
                const float s = __frcp_rn(1.f + __expf(-x));       // sigmoid(x); 1 - s saturates like the reference's fp32
                const float a1 = c_g * (1.f - s);                   // = -lr * g
                if (LOSS && sl == 0) loss_local += (x < -15.f) ? -x : -__logf(s);
                float *pu = U + (int64_t)tu * LD + sl * 4;
                float *pi = VD + (int64_t)ti * LD + sl * 4;
                float *pj = VD + (int64_t)tj * LD + sl * 4;
#pragma unroll
                for (int k = 0; k < CPL; ++k) {
                    const float4 ru = r.u[k], ri = r.i[k], rj = r.j[k];
                    float4 du, di, dj;
                    du.x = fmaf(a1, ri.x - rj.x, c_r * ru.x); du.y = fmaf(a1, ri.y - rj.y, c_r * ru.y);
                    du.z = fmaf(a1, ri.z - rj.z, c_r * ru.z); du.w = fmaf(a1, ri.w - rj.w, c_r * ru.w);
                    di.x = fmaf(a1, ru.x, c_r * ri.x); di.y = fmaf(a1, ru.y, c_r * ri.y);
                    di.z = fmaf(a1, ru.z, c_r * ri.z); di.w = fmaf(a1, ru.w, c_r * ri.w);
                    dj.x = fmaf(-a1, ru.x, c_r * rj.x); dj.y = fmaf(-a1, ru.y, c_r * rj.y);
                    dj.z = fmaf(-a1, ru.z, c_r * rj.z); dj.w = fmaf(-a1, ru.w, c_r * rj.w);
                    if (UNIQ) st4(pu + k * G * 4, make_float4(ru.x + du.x, ru.y + du.y, ru.z + du.z, ru.w + du.w));
                    else red4(pu + k * G * 4, du);
                    red4(pi + k * G * 4, di);
                    red4(pj + k * G * 4, dj);
                }
            }
        };
#pragma unroll
        for (int k = 0; k < PF; ++k) load(set[k], k);
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            if (it + PF < ITERS) load(set[(it + PF) % (PF + 1)], it + PF);
            compute(set[it % (PF + 1)], it);
        }
        c_next = __shfl_sync(0xffffffffu, c_next, 0);
    }
    if (LOSS) {
        loss_local = group_sum<32>(loss_local);
        if (lane == 0 && loss_local != 0.f) atomicAdd(p.a.loss_sum, (double)loss_local);
    }
}

// ---------------------------------------------------------------------------
// Fast path with a deep asynchronous gather: rows travel global -> shared with cp.async
// (LDGSTS.128, 16 B per lane per row piece) into a per-warp ring of S stages, one commit
// group per triple, cp.async.wait_group S-1 before the oldest is consumed.  Every lane
// reads back exactly the 16-byte pieces it copied itself, so the ring is a per-lane
// FIFO: no warp- or CTA-level synchronisation.  S triples (S * 1.5 KB at d=128) are in
// flight per warp at no register cost - the LDG fast path above is latency-bound
// (long-scoreboard 16.7 stalls/issue, DRAM 58 %, ncu run 4) with 3 in flight.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int CPL, bool UNIQ, bool LOSS, int S>
__global__ void __launch_bounds__(256) bpr_step_async_kernel(const BprParams p) {
    extern __shared__ __align__(16) float4 ring_all[];   // [8 warps][S][3][CPL][32] float4
    float *__restrict__ const U = p.a.U;
    float *__restrict__ const V = p.a.V;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float4 *const ring = ring_all + (size_t)wid * S * 3 * CPL * 32 + lane;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    constexpr int64_t LD = 128 * CPL;
    const float c_g = p.a.lr * p.invB;
    const float c_r = -p.a.lr * p.a.reg * p.invB;
    const int chunk = p.chunk, B = p.a.B;
    const int64_t n_chunks = p.n_chunks;
    b200rec_bpr_args a = p.a;
    float loss_local = 0.f;

    for (int64_t c = warp_global; c < n_chunks; c += n_warps) {
        const int64_t t_lane = c * chunk + lane;
        bool valid = (lane < chunk) && (t_lane < B);
        int u, i, j;
        fetch_triple(a, t_lane, valid, u, i, j);
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);

        auto issue = [&](int it) {
            if (it < chunk && ((vmask >> it) & 1u)) {
                const int tu = __shfl_sync(0xffffffffu, u, it);
                const int ti = __shfl_sync(0xffffffffu, i, it);
                const int tj = __shfl_sync(0xffffffffu, j, it);
                float4 *slot = ring + (size_t)(it % S) * 3 * CPL * 32;
                const float *pu = U + (int64_t)tu * LD + lane * 4;
                const float *pi = V + (int64_t)ti * LD + lane * 4;
                const float *pj = V + (int64_t)tj * LD + lane * 4;
#pragma unroll
                for (int k = 0; k < CPL; ++k) {
                    cp_async16(slot + (0 * CPL + k) * 32, pu + 128 * k);
                    cp_async16(slot + (1 * CPL + k) * 32, pi + 128 * k);
                    cp_async16(slot + (2 * CPL + k) * 32, pj + 128 * k);
                }
            }
            cp_async_commit();   // one (possibly empty) group per triple slot keeps the group count uniform
        };
#pragma unroll
        for (int it = 0; it < S; ++it) issue(it);
        for (int it = 0; it < chunk; ++it) {
            cp_async_wait<S - 1>();   // groups up to `it` have landed
            if ((vmask >> it) & 1u) {
                const int tu = __shfl_sync(0xffffffffu, u, it);
                const int ti = __shfl_sync(0xffffffffu, i, it);
                const int tj = __shfl_sync(0xffffffffu, j, it);
                const float4 *slot = ring + (size_t)(it % S) * 3 * CPL * 32;
                float4 ru[CPL], ri[CPL], rj[CPL], df[CPL];
                float part = 0.f;
#pragma unroll
                for (int k = 0; k < CPL; ++k) {
                    ru[k] = slot[(0 * CPL + k) * 32]; ri[k] = slot[(1 * CPL + k) * 32]; rj[k] = slot[(2 * CPL + k) * 32];
                    df[k] = make_float4(ri[k].x - rj[k].x, ri[k].y - rj[k].y, ri[k].z - rj[k].z, ri[k].w - rj[k].w);
                    part = fmaf(ru[k].x, df[k].x, part); part = fmaf(ru[k].y, df[k].y, part);
                    part = fmaf(ru[k].z, df[k].z, part); part = fmaf(ru[k].w, df[k].w, part);
                }
                const float x = group_sum<32>(part);   // consumes every lane's shared-memory reads of this slot
                issue(it + S);                          // refill the slot just freed
                const float s = __frcp_rn(1.f + __expf(-x));
                const float a1 = c_g * (1.f - s);
                if (LOSS) loss_local += (x < -15.f) ? -x : -__logf(s);
                float *pu = U + (int64_t)tu * LD + lane * 4;
                float *pi = V + (int64_t)ti * LD + lane * 4;
                float *pj = V + (int64_t)tj * LD + lane * 4;
#pragma unroll
                for (int k = 0; k < CPL; ++k) {
                    float4 du, di, dj;
                    du.x = fmaf(a1, df[k].x, c_r * ru[k].x); du.y = fmaf(a1, df[k].y, c_r * ru[k].y);
                    du.z = fmaf(a1, df[k].z, c_r * ru[k].z); du.w = fmaf(a1, df[k].w, c_r * ru[k].w);
                    di.x = fmaf(a1, ru[k].x, c_r * ri[k].x); di.y = fmaf(a1, ru[k].y, c_r * ri[k].y);
                    di.z = fmaf(a1, ru[k].z, c_r * ri[k].z); di.w = fmaf(a1, ru[k].w, c_r * ri[k].w);
                    dj.x = fmaf(-a1, ru[k].x, c_r * rj[k].x); dj.y = fmaf(-a1, ru[k].y, c_r * rj[k].y);
                    dj.z = fmaf(-a1, ru[k].z, c_r * rj[k].z); dj.w = fmaf(-a1, ru[k].w, c_r * rj[k].w);
                    if (UNIQ) st4(pu + 128 * k, make_float4(ru[k].x + du.x, ru[k].y + du.y, ru[k].z + du.z, ru[k].w + du.w));
                    else red4(pu + 128 * k, du);
                    red4(pi + 128 * k, di);
                    red4(pj + 128 * k, dj);
                }
            } else {
                issue(it + S);
            }
        }
        cp_async_wait<0>();
    }
    if (LOSS) {
        if (lane == 0 && loss_local != 0.f) atomicAdd(p.a.loss_sum, (double)loss_local);
    }
}

// ---------------------------------------------------------------------------
// second phase of the exact step, forward, dense optimisers
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bpr_apply_kernel(float *U, float *V, int ld, const int32_t *users,
                                                        const int32_t *pos, const int32_t *neg, int B,
                                                        const float *stage) {
    const int d4 = ld >> 2;
    const int64_t total = (int64_t)B * 3 * d4;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(e % d4);
        const int64_t rowi = e / d4;
        const int r = (int)(rowi % 3);
        const int64_t t = rowi / 3;
        float *dst = (r == 0) ? U + (int64_t)users[t] * ld : V + (int64_t)((r == 1) ? pos[t] : neg[t]) * ld;
        red4(dst + q * 4, ld4(stage + rowi * ld + q * 4));
    }
}

// sampling only (same draws as the fused kernel for the same (seed, step, t))
__global__ void __launch_bounds__(256) sample_triples_kernel(const b200rec_bpr_args a) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < a.B; t += (int64_t)gridDim.x * blockDim.x) {
        bool valid = true;
        int u, i, j;
        fetch_triple(a, t, valid, u, i, j);
    }
}

// W[ids[t]] += delta[t]   (user-gradient apply after the exchange; ids unique per call or not: atomics)
__global__ void __launch_bounds__(256) rows_add_kernel(float *W, int ld, const int32_t *ids, int n, const float *delta,
                                                       int ldd, float scale) {
    const int d4 = ld >> 2;
    const int64_t total = (int64_t)n * d4;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(e % d4);
        const int64_t t = e / d4;
        float4 v = ld4(delta + t * ldd + q * 4);
        v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
        if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) red4(W + (int64_t)ids[t] * ld + q * 4, v);
    }
}

__global__ void __launch_bounds__(256) mf_forward_kernel(const float *U, const float *V, int ld, int d,
                                                         const int32_t *users, const int32_t *items, int n,
                                                         float *out) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int d4 = ld >> 2;
    for (int64_t t = w; t < n; t += nw) {
        const float *pu = U + (int64_t)users[t] * ld, *pv = V + (int64_t)items[t] * ld;
        float acc = 0.f;
        for (int q = lane; q < d4; q += 32) {
            const float4 x = ld4(pu + q * 4), y = ld4(pv + q * 4);
            acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc); acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
        }
        acc = group_sum<32>(acc);
        if (lane == 0) out[t] = acc;
    }
}

__global__ void __launch_bounds__(256) sgd_dense_kernel(float *p, const float *g, int64_t n, float lr) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
        p[e] = p[e] - lr * g[e];
}

// torch.optim.Adam single-tensor update (no amsgrad, no weight decay, not maximize):
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
__global__ void __launch_bounds__(256) adam_dense_kernel(float *p, const float *g, float *m, float *v, int64_t n,
                                                         float b1, float b2, float eps, float step_size,
                                                         float bc2_sqrt) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const float gg = g[e];
        const float mm = m[e] * b1 + (1.f - b1) * gg;              // lerp form used by torch: m + (g-m)*(1-b1)
        const float vv = v[e] * b2 + (1.f - b2) * gg * gg;
        m[e] = mm; v[e] = vv;
        const float denom = sqrtf(vv) / bc2_sqrt + eps;
        p[e] = p[e] - step_size * (mm / denom);
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
template <int G, int CPL>
static int launch_bpr(BprParams &p, cudaStream_t s) {
    const bool tma = (p.a.flags & B200REC_F_TMA_GATHER) != 0;
    {   // stages per warp: as many as fit in ~96 KB per CTA (2 CTAs/SM), between 2 and kTmaMaxStages
        const size_t per_stage = (size_t)kTmaWarps * p.a.ld * 4 * 3 * (32 / G);
        int st = (int)((96 * 1024) / per_stage);
        p.stages = st < 2 ? 2 : (st > kTmaMaxStages ? kTmaMaxStages : st);
    }
    // B200REC_SM_RESERVE=n leaves n SMs' worth of CTA slots free (room for a concurrent NCCL kernel)
    static const int reserve = getenv("B200REC_SM_RESERVE") ? atoi(getenv("B200REC_SM_RESERVE")) : 0;
    const int sms = sm_count() - ((reserve > 0 && reserve < sm_count()) ? reserve : 0);
    constexpr int TPW = 32 / G;
#define B200_LAUNCH_SINK(SINKV)                                                                              \
    if (tma) {                                                                                               \
        auto kern = bpr_step_tma_kernel<G, CPL, SINKV>;                                                      \
        const size_t smem = (size_t)kTmaWarps * (kTmaMaxStages * 8 + (size_t)p.stages * p.a.ld * 4 * 3 * TPW);  \
        B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
        int occ = 0;                                                                                         \
        B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kTmaWarps * 32, smem));          \
        if (occ < 1) occ = 1;                                                                                \
        int64_t need = (p.n_chunks + kTmaWarps - 1) / kTmaWarps;                                             \
        int grid = (int)(need < (int64_t)sms * occ ? need : (int64_t)sms * occ);                             \
        if (grid < 1) grid = 1;                                                                              \
        kern<<<grid, kTmaWarps * 32, smem, s>>>(p);                                                          \
    } else {                                                                                                 \
        auto kern = bpr_step_ldg_kernel<G, CPL, SINKV>;                                                      \
        int occ = 0;                                                                                         \
        B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0));                        \
        if (occ < 1) occ = 1;                                                                                \
        int64_t need = (p.n_chunks + 7) / 8;                                                                 \
        int grid = (int)(need < (int64_t)sms * occ ? need : (int64_t)sms * occ);                             \
        if (grid < 1) grid = 1;                                                                              \
        kern<<<grid, 256, 0, s>>>(p);                                                                        \
    }
    switch (p.a.sink) {
        case B200REC_SINK_UPDATE: { B200_LAUNCH_SINK(B200REC_SINK_UPDATE) } break;
        case B200REC_SINK_STAGE: { B200_LAUNCH_SINK(B200REC_SINK_STAGE) } break;
        case B200REC_SINK_GRAD: { B200_LAUNCH_SINK(B200REC_SINK_GRAD) } break;
        default: { B200_LAUNCH_SINK(B200REC_SINK_NONE) } break;
    }
#undef B200_LAUNCH_SINK
    B200_LAUNCH_CHECK();
    return B200REC_OK;
}

}  // namespace b200

using namespace b200;

// name of the kernel the last b200rec_bpr_step of this process dispatched to (bench.py ties its roofline.traffic figure,
// an ncu capture of ONE named kernel, to what actually ran)
static char g_last_step_kernel[96] = "";
extern "C" const char *b200rec_last_step_kernel(void) { return g_last_step_kernel; }
#define B200_NOTE_KERNEL(...) snprintf(g_last_step_kernel, sizeof(g_last_step_kernel), __VA_ARGS__)

extern "C" int b200rec_bpr_step(const b200rec_bpr_args *args, void *stream) {
    B200_REQUIRE(args != nullptr, B200REC_EINVAL, "bpr_step: args is NULL");
    const b200rec_bpr_args &a = *args;
    if (a.B == 0) return B200REC_OK;
    B200_REQUIRE(a.U && a.V && a.users, B200REC_EINVAL, "bpr_step: U, V and users are required");
    B200_REQUIRE(a.B >= 0 && a.d >= 1 && a.ld >= a.d && (a.ld % 4) == 0 && a.ld <= 512, B200REC_EINVAL,
                 "bpr_step: need 1 <= d <= ld <= 512, ld %% 4 == 0 (d=%d ld=%d)", a.d, a.ld);
    B200_REQUIRE(((uintptr_t)a.U % 16) == 0 && ((uintptr_t)a.V % 16) == 0, B200REC_EINVAL,
                 "bpr_step: tables must be 16-byte aligned");
    B200_REQUIRE((a.pos && a.neg) || (a.csr_indptr && a.csr_indices && a.num_items > 0), B200REC_EINVAL,
                 "bpr_step: on-device sampling needs the CSR of positives");
    B200_REQUIRE(a.sink >= 0 && a.sink <= 3, B200REC_EINVAL, "bpr_step: bad sink %d", a.sink);
    B200_REQUIRE(a.sink != B200REC_SINK_STAGE || a.stage, B200REC_EINVAL, "bpr_step: SINK_STAGE needs stage");
    B200_REQUIRE(a.sink != B200REC_SINK_GRAD || (a.gU && a.gV), B200REC_EINVAL, "bpr_step: SINK_GRAD needs gU,gV");
    B200_REQUIRE(!(a.flags & B200REC_F_ITEM_DELTA) || a.gV, B200REC_EINVAL, "bpr_step: F_ITEM_DELTA needs gV");
    B200_REQUIRE(!(a.flags & B200REC_F_ITEM_DELTA_BF16) || ((a.flags & B200REC_F_ITEM_DELTA) && a.sink == B200REC_SINK_UPDATE),
                 B200REC_EINVAL, "bpr_step: F_ITEM_DELTA_BF16 refines F_ITEM_DELTA (SINK_UPDATE)");
    B200_REQUIRE(a.item_hi >= a.item_lo && a.item_lo >= 0, B200REC_EINVAL, "bpr_step: bad item shard range");
    B200_REQUIRE(a.item_hi == a.item_lo || (a.pos == nullptr && a.neg == nullptr) || a.item_lo == 0, B200REC_EINVAL,
                 "bpr_step: item-sharded mode samples its own triples (pos/neg must be NULL)");
    if (a.B == 0) return B200REC_OK;

    const int d4 = a.ld / 4;
    int G = 1;
    while (G < d4 && G < 32) G <<= 1;
    const int CPL = (d4 + G - 1) / G;
    const int TPW = 32 / G;
    BprParams p;
    p.a = a;
    p.work = nullptr;
    p.invB = a.inv_batch > 0.f ? a.inv_batch : 1.0f / (float)a.B;
    // chunk: as large as 32 triples, shrunk (to a multiple of TPW) until every warp slot has work
    const int64_t slots = (int64_t)sm_count() * 48;
    int chunk = 32;
    while (chunk > TPW && chunk > 4 && ((int64_t)a.B + chunk - 1) / chunk < slots) chunk >>= 1;
    if (chunk < TPW) chunk = TPW;
    p.chunk = chunk;
    p.n_chunks = ((int64_t)a.B + chunk - 1) / chunk;
    cudaStream_t s = (cudaStream_t)stream;
    // lean fast path (see bpr_step_fast_kernel): the plain single-device fused update on full-warp rows
    if (a.sink == B200REC_SINK_UPDATE && G == 32 && d4 == 32 * CPL && CPL <= 2 &&
        !(a.flags & (B200REC_F_TMA_GATHER | B200REC_F_ITEM_DELTA_BF16 | B200REC_F_GENERIC)) && !a.udelta &&
        a.item_hi == a.item_lo && !a.x_out &&
        !((a.flags & B200REC_F_ITEM_DELTA) && (a.flags & B200REC_F_ASYNC_GATHER))) {
        const bool uniq = (a.flags & B200REC_F_USERS_UNIQUE) != 0, loss = a.loss_sum != nullptr;
        const bool idelta = (a.flags & B200REC_F_ITEM_DELTA) != 0;
        if (a.flags & B200REC_F_ASYNC_GATHER) {   // deep cp.async ring (bpr_step_async_kernel)
            const size_t smem = (size_t)8 * 8 * 3 * 512;   // 8 warps x (S*CPL = 8) x 3 rows x 512 B = 96 KB
#define B200_ASYNC(C, Q, L, SS)                                                                                \
    {                                                                                                          \
        auto kern = bpr_step_async_kernel<C, Q, L, SS>;                                                        \
        B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
        int occ = 0;                                                                                           \
        B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem));                       \
        if (occ < 1) occ = 1;                                                                                  \
        const int64_t need_a = (p.n_chunks + 7) / 8, cap_a = (int64_t)sm_count() * occ;                        \
        kern<<<(int)(need_a < cap_a ? need_a : cap_a), 256, smem, s>>>(p);                                     \
    }
            if (CPL == 1) {
                if (uniq) { if (loss) B200_ASYNC(1, true, true, 8) else B200_ASYNC(1, true, false, 8) }
                else { if (loss) B200_ASYNC(1, false, true, 8) else B200_ASYNC(1, false, false, 8) }
            } else {
                if (uniq) { if (loss) B200_ASYNC(2, true, true, 4) else B200_ASYNC(2, true, false, 4) }
                else { if (loss) B200_ASYNC(2, false, true, 4) else B200_ASYNC(2, false, false, 4) }
            }
#undef B200_ASYNC
            B200_NOTE_KERNEL("bpr_step_async_kernel<%d,%d,%d>", CPL, (int)uniq, (int)loss);
            B200_LAUNCH_CHECK();
            return B200REC_OK;
        }
        // dynamic chunk distribution (B200REC_FAST_DYNAMIC=0 to disable): a ring of counters kept per device, so
        // back-to-back launches on different streams never share one
        {
            static unsigned int *ring[64] = {nullptr};
            static unsigned int slot[64] = {0};
            static const bool dyn = !(getenv("B200REC_FAST_DYNAMIC") && atoi(getenv("B200REC_FAST_DYNAMIC")) == 0);
            int dev = 0;
            p.work = nullptr;
            if (dyn && cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) {
                if (!ring[dev]) B200_CUDA(cudaMalloc(&ring[dev], 256 * sizeof(unsigned int)));
                p.work = ring[dev] + (slot[dev]++ & 255u);
                B200_CUDA(cudaMemsetAsync(p.work, 0, sizeof(unsigned int), s));
            }
        }
#ifdef B200REC_ABLATE
        {
            const int abl = getenv("B200REC_ABL") ? atoi(getenv("B200REC_ABL")) : 0;
            B200_CUDA(cudaMemcpyToSymbolAsync(c_abl, &abl, sizeof(int), 0, cudaMemcpyHostToDevice, s));
        }
#endif
        // group kernel (default for ld = 128; B200REC_STEP_VARIANT="G,PF" selects another instantiation, "0,0" the
        // warp-per-row kernel below)
        {
            const char *var = getenv("B200REC_STEP_VARIANT");
            int vg = 8, vpf = 0;      // measured best at cfg2 (profiles/r02_bpr_ablation.md); "0,0" = warp-per-row kernel
            if (var && sscanf(var, "%d,%d", &vg, &vpf) != 2) { vg = 8; vpf = 0; }
            if (vg > 0 && p.work && CPL == 1 && !(a.flags & B200REC_F_L2_HINTS)) {
                p.chunk = 32;
                p.n_chunks = ((int64_t)a.B + 31) / 32;
#define B200_GROUP(GG, PP)                                                                                   \
    if (vg == GG && vpf == PP) {                                                                             \
        auto launch = [&](auto kern) -> int {                                                                \
            int occ = 0;                                                                                     \
            B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0));                    \
            if (occ < 1) occ = 1;                                                                            \
            const int64_t need_g = (p.n_chunks + 7) / 8, cap_g = (int64_t)sm_count() * occ;                  \
            kern<<<(int)(need_g < cap_g ? need_g : cap_g), 256, 0, s>>>(p);                                  \
            return B200REC_OK;                                                                               \
        };                                                                                                   \
        int rc_g;                                                                                            \
        if (uniq) { if (loss) { if (idelta) rc_g = launch(bpr_step_group_kernel<GG, 32, PP, true, true, true>); else rc_g = launch(bpr_step_group_kernel<GG, 32, PP, true, true, false>); } \
                    else { if (idelta) rc_g = launch(bpr_step_group_kernel<GG, 32, PP, true, false, true>); else rc_g = launch(bpr_step_group_kernel<GG, 32, PP, true, false, false>); } } \
        else { if (loss) { if (idelta) rc_g = launch(bpr_step_group_kernel<GG, 32, PP, false, true, true>); else rc_g = launch(bpr_step_group_kernel<GG, 32, PP, false, true, false>); } \
               else { if (idelta) rc_g = launch(bpr_step_group_kernel<GG, 32, PP, false, false, true>); else rc_g = launch(bpr_step_group_kernel<GG, 32, PP, false, false, false>); } } \
        if (rc_g) return rc_g;                                                                               \
        B200_NOTE_KERNEL("bpr_step_group_kernel<%d,32,%d,%d,%d,%d>", GG, PP, (int)uniq, (int)loss, (int)idelta); \
        B200_LAUNCH_CHECK();                                                                                 \
        return B200REC_OK;                                                                                   \
    }
                B200_GROUP(8, 0) B200_GROUP(8, 1) B200_GROUP(16, 1)
#undef B200_GROUP
            }
        }
        const int64_t need = (p.n_chunks + 7) / 8;
        const int64_t cap = (int64_t)sm_count() * (CPL == 1 ? B200REC_FAST_MINB : 2);
        const int grid = (int)(need < cap ? need : cap);
#define B200_FAST(C, Q, L)                                                              \
    {                                                                                   \
        if (idelta) bpr_step_fast_kernel<C, Q, L, true><<<grid, 256, 0, s>>>(p);        \
        else bpr_step_fast_kernel<C, Q, L, false><<<grid, 256, 0, s>>>(p);              \
    }
        if (CPL == 1) {
            if (uniq) { if (loss) B200_FAST(1, true, true) else B200_FAST(1, true, false) }
            else { if (loss) B200_FAST(1, false, true) else B200_FAST(1, false, false) }
        } else {
            if (uniq) { if (loss) B200_FAST(2, true, true) else B200_FAST(2, true, false) }
            else { if (loss) B200_FAST(2, false, true) else B200_FAST(2, false, false) }
        }
#undef B200_FAST
        B200_NOTE_KERNEL("bpr_step_fast_kernel<%d,%d,%d,%d>", CPL, (int)uniq, (int)loss, (int)idelta);
        B200_LAUNCH_CHECK();
        return B200REC_OK;
    }
    B200_NOTE_KERNEL("%s<%d,%d,sink%d>", (a.flags & B200REC_F_TMA_GATHER) ? "bpr_step_tma_kernel" : "bpr_step_ldg_kernel", G, CPL,
                     a.sink);
    switch (G) {
        case 1: return launch_bpr<1, 1>(p, s);
        case 2: return launch_bpr<2, 1>(p, s);
        case 4: return launch_bpr<4, 1>(p, s);
        case 8: return launch_bpr<8, 1>(p, s);
        case 16: return launch_bpr<16, 1>(p, s);
        default:
            switch (CPL) {
                case 1: return launch_bpr<32, 1>(p, s);
                case 2: return launch_bpr<32, 2>(p, s);
                case 3: return launch_bpr<32, 3>(p, s);
                default: return launch_bpr<32, 4>(p, s);
            }
    }
}

extern "C" int b200rec_bpr_apply(float *U, float *V, int ld, const int32_t *users, const int32_t *pos,
                                 const int32_t *neg, int B, const float *stage, void *stream) {
    B200_REQUIRE(U && V && users && pos && neg && stage, B200REC_EINVAL, "bpr_apply: null argument");
    B200_REQUIRE(ld > 0 && ld % 4 == 0, B200REC_EINVAL, "bpr_apply: ld %% 4 != 0");
    if (B <= 0) return B200REC_OK;
    const int64_t total = (int64_t)B * 3 * (ld / 4);
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    bpr_apply_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(U, V, ld, users, pos, neg,
                                                                                           B, stage);
    B200_LAUNCH_CHECK();
    return B200REC_OK;
}

extern "C" int b200rec_mf_forward(const float *U, const float *V, int ld, int d, const int32_t *users,
                                  const int32_t *items, int n, float *out, void *stream) {
    B200_REQUIRE(U && V && users && items && out, B200REC_EINVAL, "mf_forward: null argument");
    B200_REQUIRE(d >= 1 && ld >= d && ld % 4 == 0, B200REC_EINVAL, "mf_forward: bad d/ld");
    if (n <= 0) return B200REC_OK;
    int64_t blocks = ((int64_t)n + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 8;
    mf_forward_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(U, V, ld, d, users, items,
                                                                                            n, out);
    B200_LAUNCH_CHECK();
    return B200REC_OK;
}

extern "C" int b200rec_sgd_dense(float *param, const float *grad, int64_t n, float lr, void *stream) {
    B200_REQUIRE(param && grad, B200REC_EINVAL, "sgd_dense: null argument");
    if (n <= 0) return B200REC_OK;
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    sgd_dense_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(param, grad, n, lr);
    B200_LAUNCH_CHECK();
    return B200REC_OK;
}

// ---- "update in place, exchange the difference" (user-sharded multi-GPU layout) ---------------------------------
// d_own = d_wire = W - snapshot  (what this rank's step did to its replica; two copies: one is all-reduced in place,
// the other is subtracted again when the sum comes back)
__global__ void __launch_bounds__(256) delta_diff_kernel(const float4 *__restrict__ W, const float4 *__restrict__ snap,
                                                         float4 *__restrict__ d_wire, float4 *__restrict__ d_own,
                                                         int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 w = W[i], s = snap[i];
        const float4 d = make_float4(w.x - s.x, w.y - s.y, w.z - s.z, w.w - s.w);
        d_wire[i] = d;
        d_own[i] = d;
    }
}
// W += d_sum - d_own  (the other ranks' contribution of that step)
__global__ void __launch_bounds__(256) delta_apply_kernel(float4 *__restrict__ W, const float4 *__restrict__ d_sum,
                                                          const float4 *__restrict__ d_own, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 w = W[i];
        const float4 a = d_sum[i], b = d_own[i];
        w.x += a.x - b.x; w.y += a.y - b.y; w.z += a.z - b.z; w.w += a.w - b.w;
        W[i] = w;
    }
}

// W += d; d = 0   (replicated head rows of the P2P layout: apply the all-reduced delta and clear the buffer in one pass)
__global__ void __launch_bounds__(256) add_clear_kernel(float4 *__restrict__ W, float4 *__restrict__ d, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 w = W[i];
        const float4 a = d[i];
        w.x += a.x; w.y += a.y; w.z += a.z; w.w += a.w;
        W[i] = w;
        d[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// W = snapshot + scale * d   (replicated head rows: combine the all-reduced per-rank differences; scale = 1 is the
// data-parallel sum, scale = 1/ranks the per-step parameter average)
__global__ void __launch_bounds__(256) snap_apply_kernel(float4 *__restrict__ W, const float4 *__restrict__ snap,
                                                         const float4 *__restrict__ d, float scale, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 s = snap[i], a = d[i];
        W[i] = make_float4(fmaf(scale, a.x, s.x), fmaf(scale, a.y, s.y), fmaf(scale, a.z, s.z), fmaf(scale, a.w, s.w));
    }
}

extern "C" int b200rec_snap_apply(float *W, const float *snapshot, const float *d_sum, float scale, int64_t n, void *stream) {
    B200_REQUIRE(W && snapshot && d_sum, B200REC_EINVAL, "snap_apply: null argument");
    B200_REQUIRE(n % 4 == 0, B200REC_EINVAL, "snap_apply: n must be a multiple of 4 (padded tables)");
    if (n <= 0) return B200REC_OK;
    const int64_t n4 = n / 4, blocks = (n4 + 255) / 256, cap = (int64_t)sm_count() * 16;
    snap_apply_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<float4 *>(W), reinterpret_cast<const float4 *>(snapshot), reinterpret_cast<const float4 *>(d_sum), scale, n4);
    B200_LAUNCH_CHECK();
    return B200REC_OK;
}

extern "C" int b200rec_add_clear(float *W, float *delta, int64_t n, void *stream) {
    B200_REQUIRE(W && delta, B200REC_EINVAL, "add_clear: null argument");
    B200_REQUIRE(n % 4 == 0, B200REC_EINVAL, "add_clear: n must be a multiple of 4 (padded tables)");
    if (n <= 0) return B200REC_OK;
    const int64_t n4 = n / 4, blocks = (n4 + 255) / 256, cap = (int64_t)sm_count() * 16;
    add_clear_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<float4 *>(W), reinterpret_cast<float4 *>(delta), n4);
    B200_LAUNCH_CHECK();
    return B200REC_OK;
}

extern "C" int b200rec_delta_diff(const float *W, const float *snapshot, float *d_wire, float *d_own, int64_t n,
                                  void *stream) {
    B200_REQUIRE(W && snapshot && d_wire && d_own, B200REC_EINVAL, "delta_diff: null argument");
    B200_REQUIRE(n % 4 == 0, B200REC_EINVAL, "delta_diff: n must be a multiple of 4 (padded tables)");
    if (n <= 0) return B200REC_OK;
    const int64_t n4 = n / 4, blocks = (n4 + 255) / 256, cap = (int64_t)sm_count() * 16;
    delta_diff_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(W), reinterpret_cast<const float4 *>(snapshot),
        reinterpret_cast<float4 *>(d_wire), reinterpret_cast<float4 *>(d_own), n4);
    B200_LAUNCH_CHECK();
    return B200REC_OK;
}

extern "C" int b200rec_delta_apply(float *W, const float *d_sum, const float *d_own, int64_t n, void *stream) {
    B200_REQUIRE(W && d_sum && d_own, B200REC_EINVAL, "delta_apply: null argument");
    B200_REQUIRE(n % 4 == 0, B200REC_EINVAL, "delta_apply: n must be a multiple of 4 (padded tables)");
    if (n <= 0) return B200REC_OK;
    const int64_t n4 = n / 4, blocks = (n4 + 255) / 256, cap = (int64_t)sm_count() * 16;
    delta_apply_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<float4 *>(W), reinterpret_cast<const float4 *>(d_sum), reinterpret_cast<const float4 *>(d_own), n4);
    B200_LAUNCH_CHECK();
    return B200REC_OK;
}

extern "C" int b200rec_adam_dense(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n,
                                  float lr, float beta1, float beta2, float eps, int step, void *stream) {
    B200_REQUIRE(param && grad && exp_avg && exp_avg_sq, B200REC_EINVAL, "adam_dense: null argument");
    B200_REQUIRE(step >= 1, B200REC_EINVAL, "adam_dense: step counts from 1");
    if (n <= 0) return B200REC_OK;
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    adam_dense_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(
        param, grad, exp_avg, exp_avg_sq, n, beta1, beta2, eps, (float)((double)lr / bc1), (float)sqrt(bc2));
    B200_LAUNCH_CHECK();
    return B200REC_OK;
}

extern "C" int b200rec_sample_triples(const int32_t *users, int B, const int64_t *csr_indptr,
                                      const int32_t *csr_indices, int num_items, uint64_t seed, uint64_t step,
                                      int32_t *out_pos, int32_t *out_neg, void *stream) {
    B200_REQUIRE(users && csr_indptr && csr_indices && out_pos && out_neg && num_items > 0, B200REC_EINVAL,
                 "sample_triples: null argument");
    if (B <= 0) return B200REC_OK;
    b200rec_bpr_args a;
    memset(&a, 0, sizeof(a));
    a.users = users; a.B = B; a.csr_indptr = csr_indptr; a.csr_indices = csr_indices; a.num_items = num_items;
    a.seed = seed; a.step = step; a.out_pos = out_pos; a.out_neg = out_neg;
    int64_t blocks = ((int64_t)B + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    sample_triples_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(a);
    B200_LAUNCH_CHECK();
    return B200REC_OK;
}

extern "C" int b200rec_rows_add(float *W, int ld, const int32_t *ids, int n, const float *delta, int ld_delta,
                                float scale, void *stream) {
    B200_REQUIRE(W && ids && delta, B200REC_EINVAL, "rows_add: null argument");
    B200_REQUIRE(ld > 0 && ld % 4 == 0 && ld_delta >= ld && ld_delta % 4 == 0, B200REC_EINVAL, "rows_add: bad ld");
    if (n <= 0) return B200REC_OK;
    const int64_t total = (int64_t)n * (ld / 4);
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    rows_add_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(W, ld, ids, n, delta,
                                                                                          ld_delta, scale);
    B200_LAUNCH_CHECK();
    return B200REC_OK;
}

// ---------------------------------------------------------------------------
// Row-wise (lazy) Adam - SURVEY section 8(f) rank 1.  torch.optim.SparseAdam semantics: only the rows present in the
// batch move; for such a row (gradient g summed over its occurrences, all from pre-step weights):
//   m = m + (1-b1)(g - m);  v = v + (1-b2)(g*g - v);  w -= lr*sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps)
// The fused step first accumulates g into dense scratch rows (SINK_GRAD); this kernel then walks the batch ids,
// lets exactly one occurrence of each row claim it (atomicExch on a per-row stamp holding the step number),
// applies the update and zeroes the scratch row again (so the scratch never needs a dense memset).
// ---------------------------------------------------------------------------
namespace b200 {
template <int G, int CPL>
__global__ void __launch_bounds__(256) adam_rows_kernel(float *__restrict__ W, float *__restrict__ g,
                                                        float *__restrict__ m, float *__restrict__ v,
                                                        int32_t *__restrict__ stamp, int ld,
                                                        const int32_t *__restrict__ ids, int n, float b1, float b2,
                                                        float eps, float step_size, int step) {
    constexpr int RPW = 32 / G;
    const int lane = threadIdx.x & 31;
    const int sl = lane % G, sg = lane / G;
    const int d4 = ld >> 2;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (sg * G));
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t tb = warp_global * RPW; tb < n; tb += n_warps * RPW) {
        const int64_t t = tb + sg;
        int id = -1;
        if (t < n) id = ids[t];
        int mine = 0;
        if (sl == 0 && id >= 0) mine = (atomicExch(stamp + id, step) != step);
        mine = __shfl_sync(gmask, mine, sg * G);
        if (!mine) continue;
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            const int q = sl + k * G;
            if (q < d4) {
                const int64_t off = (int64_t)id * ld + q * 4;
                const float4 gg = ld4(g + off);
                float4 mm = ld4(m + off), vv = ld4(v + off), ww = ld4(W + off);
                mm.x += (1.f - b1) * (gg.x - mm.x); mm.y += (1.f - b1) * (gg.y - mm.y);
                mm.z += (1.f - b1) * (gg.z - mm.z); mm.w += (1.f - b1) * (gg.w - mm.w);
                vv.x += (1.f - b2) * (gg.x * gg.x - vv.x); vv.y += (1.f - b2) * (gg.y * gg.y - vv.y);
                vv.z += (1.f - b2) * (gg.z * gg.z - vv.z); vv.w += (1.f - b2) * (gg.w * gg.w - vv.w);
                ww.x -= step_size * (mm.x / (sqrtf(vv.x) + eps)); ww.y -= step_size * (mm.y / (sqrtf(vv.y) + eps));
                ww.z -= step_size * (mm.z / (sqrtf(vv.z) + eps)); ww.w -= step_size * (mm.w / (sqrtf(vv.w) + eps));
                st4(m + off, mm); st4(v + off, vv); st4(W + off, ww);
                st4(g + off, make_float4(0.f, 0.f, 0.f, 0.f));
            }
        }
    }
}

template <int G, int CPL>
static int launch_adam_rows(float *W, float *g, float *m, float *v, int32_t *stamp, int ld, const int32_t *ids, int n,
                            float b1, float b2, float eps, float step_size, int step, cudaStream_t s) {
    constexpr int RPW = 32 / G;
    int64_t blocks = ((int64_t)n + 8 * RPW - 1) / (8 * RPW);
    const int64_t cap = (int64_t)sm_count() * 8;
    adam_rows_kernel<G, CPL><<<(int)(blocks < cap ? blocks : cap), 256, 0, s>>>(W, g, m, v, stamp, ld, ids, n, b1, b2,
                                                                                eps, step_size, step);
    B200_LAUNCH_CHECK();
    return B200REC_OK;
}
}  // namespace b200

extern "C" int b200rec_adam_rows(float *W, float *grad, float *exp_avg, float *exp_avg_sq, int32_t *stamp, int ld,
                                 const int32_t *ids, int n, float lr, float beta1, float beta2, float eps, int step,
                                 void *stream) {
    B200_REQUIRE(W && grad && exp_avg && exp_avg_sq && stamp && ids, B200REC_EINVAL, "adam_rows: null argument");
    B200_REQUIRE(ld > 0 && ld % 4 == 0 && ld <= 512 && step >= 1, B200REC_EINVAL, "adam_rows: bad ld/step");
    if (n <= 0) return B200REC_OK;
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    const float step_size = (float)((double)lr * sqrt(bc2) / bc1);
    const int d4 = ld / 4;
    int G = 1;
    while (G < d4 && G < 32) G <<= 1;
    const int CPL = (d4 + G - 1) / G;
    cudaStream_t s = (cudaStream_t)stream;
#define B200_AR(GG, CC) return launch_adam_rows<GG, CC>(W, grad, exp_avg, exp_avg_sq, stamp, ld, ids, n, beta1, beta2, eps, step_size, step, s)
    switch (G) {
        case 1: B200_AR(1, 1);
        case 2: B200_AR(2, 1);
        case 4: B200_AR(4, 1);
        case 8: B200_AR(8, 1);
        case 16: B200_AR(16, 1);
        default:
            switch (CPL) {
                case 1: B200_AR(32, 1);
                case 2: B200_AR(32, 2);
                case 3: B200_AR(32, 3);
                default: B200_AR(32, 4);
            }
    }
#undef B200_AR
}

// V += float(dV) for the bf16 item-delta exchange buffer of the user-sharded layout
namespace b200 {
__global__ void __launch_bounds__(256) add_bf16_kernel(float *__restrict__ W, const __nv_bfloat16 *__restrict__ delta,
                                                       int64_t n4) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (int64_t)gridDim.x * blockDim.x) {
        const uint2 raw = *reinterpret_cast<const uint2 *>(delta + e * 4);
        const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162 *>(&raw.x);
        const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162 *>(&raw.y);
        float4 w = ld4(W + e * 4);
        w.x += __low2float(lo); w.y += __high2float(lo); w.z += __low2float(hi); w.w += __high2float(hi);
        st4(W + e * 4, w);
    }
}
}  // namespace b200

extern "C" int b200rec_add_bf16(float *W, const void *delta_bf16, int64_t n, void *stream) {
    B200_REQUIRE(W && delta_bf16 && n % 4 == 0, B200REC_EINVAL, "add_bf16: null argument or n %% 4 != 0");
    if (n <= 0) return B200REC_OK;
    const int64_t n4 = n / 4;
    int64_t blocks = (n4 + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    add_bf16_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(
        W, reinterpret_cast<const __nv_bfloat16 *>(delta_bf16), n4);
    B200_LAUNCH_CHECK();
    return B200REC_OK;
}
