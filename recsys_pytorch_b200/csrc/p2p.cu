// Multi-GPU BPR step with the exchange INSIDE the kernel: item table sharded by item-id range (BASELINE north_star),
// user table sharded by user-id range, user rows travel through NVSwitch peer memory (P2P loads / stores issued by the
// fused step kernel itself) instead of through a separate collective.
//
// The reference is single-device; what this replaces per step is still models/MF.py:63-68 +
// data/generators.py:168-201, exactly as bpr_step.cu does on one GPU.  Per step and rank:
//
//   route   (local)  sample (i, j) for the rank's own batch users from its CSR shard, find owner(i) in the item
//                    bounds, draw j from owner(i)'s range (i and j stay co-located - the north_star deviation from
//                    generators.py:178-189: "uniform over the owner shard's non-positives"), and append (u, i, j) to
//                    the outbox segment of that owner.  Outboxes live in the peer-mapped arena of the routing rank (CUDA VMM / IPC).
//   barrier          one tiny stream-ordered all-reduce (host side, dist.py) - routes of step s and steps <= s-1 done.
//   step    (fused)  every rank PULLS the triples routed to it from all outboxes (coalesced peer loads of the ids),
//                    loads the user row from its HOME rank's table over NVLink (512 B at d=128), the two item rows
//                    from its own shard (L2-resident: 64 MB per GPU at cfg3), computes x, g, applies the item updates
//                    with local vector atomics and writes the updated user row back to its home (peer store).
//                    NVLink traffic: 4*ld bytes per triple per direction for the (W-1)/W remote fraction - half of a
//                    replicated-table all-gather of user deltas.
//
// Fixed-triple parity mode (SURVEY 8(e) bullet 2): with given (pos, neg) the negative may live on another shard; the
// step kernel then resolves V[j] through the peer table too (peer load + peer vector atomic), so ANY triple list gives
// the single-device result up to fp32 summation order.
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "sampler.cuh"

namespace b200 {

__device__ __forceinline__ int owner_of(const int32_t *bounds, int world, int id) {
    int r = 0;
#pragma unroll 1
    for (int k = 1; k < world; ++k) r += (id >= bounds[k]);
    return r;
}

// ---------------------------------------------------------------------------------------------------------------
// route: sample + bucket by owner of the positive.  One triple per thread; a CTA aggregates its appends so that the
// outbox counters see W atomics per CTA tile instead of one per triple.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kRouteThreads = 256;
constexpr int kRouteTilesPerCta = 4;   // consecutive tiles handled by one CTA iteration (fewer counter atomics)

__global__ void __launch_bounds__(kRouteThreads) p2p_route_kernel(const b200rec_p2p_route_args a) {
    __shared__ int hist[B200REC_MAX_RANKS];
    __shared__ int base[B200REC_MAX_RANKS];
    const int W = a.world;
    const int64_t tile = (int64_t)kRouteThreads * kRouteTilesPerCta;
    for (int64_t t0 = (int64_t)blockIdx.x * tile; t0 < a.B; t0 += (int64_t)gridDim.x * tile) {
        if (threadIdx.x < B200REC_MAX_RANKS) hist[threadIdx.x] = 0;
        __syncthreads();
        int u[kRouteTilesPerCta], i[kRouteTilesPerCta], j[kRouteTilesPerCta], dst[kRouteTilesPerCta], slot[kRouteTilesPerCta];
#pragma unroll
        for (int k = 0; k < kRouteTilesPerCta; ++k) {
            const int64_t t = t0 + (int64_t)k * kRouteThreads + threadIdx.x;
            dst[k] = -1; u[k] = i[k] = j[k] = 0; slot[k] = 0;
            if (t < a.B) {
                u[k] = a.users[t];
                bool valid = true;
                if (a.pos) i[k] = a.pos[t];
                if (a.neg) j[k] = a.neg[t];
                if (!a.pos || !a.neg) {
                    const int64_t lo = a.csr_indptr[u[k]], hi = a.csr_indptr[u[k] + 1];
                    const uint32_t deg = (uint32_t)(hi - lo);
                    const int32_t *row = a.csr_indices + lo;
                    if (deg == 0 && !a.pos) valid = false;
                    if (valid && !a.pos) i[k] = sample_pos(row, deg, a.seed, a.step, (uint64_t)t);
                    if (valid && !a.neg) {
                        const int o = i[k] < a.head ? a.rank : owner_of(a.item_bounds, W, i[k]);
                        const uint32_t n_lo = (uint32_t)a.item_bounds[o];
                        valid = sample_neg2(row, deg, (uint32_t)a.head, n_lo, (uint32_t)a.item_bounds[o + 1] - n_lo,
                                            a.seed, a.step, (uint64_t)t, j[k]);
                    }
                }
                if (valid) {
                    dst[k] = i[k] < a.head ? a.rank : owner_of(a.item_bounds, W, i[k]);
                    slot[k] = atomicAdd(&hist[dst[k]], 1);
                }
                if (a.dbg_pos) a.dbg_pos[t] = valid ? i[k] : -1;
                if (a.dbg_neg) a.dbg_neg[t] = valid ? j[k] : -1;
            }
        }
        __syncthreads();
        if (threadIdx.x < W) base[threadIdx.x] = hist[threadIdx.x] ? atomicAdd(a.out_cnt + threadIdx.x, hist[threadIdx.x]) : 0;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kRouteTilesPerCta; ++k) {
            if (dst[k] >= 0) {
                const int64_t o = (int64_t)dst[k] * a.cap + base[dst[k]] + slot[k];
                a.out_u[o] = u[k]; a.out_i[o] = i[k]; a.out_j[o] = j[k];
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// fused step over the triples routed to this rank
// ---------------------------------------------------------------------------------------------------------------
// Work order.  The dynamic chunk counter c walks the remote homes ONE AFTER THE OTHER (me+1, me+2, ...) and slips one
// chunk of this rank's own (NVLink-free) triples in after every W-1 remote chunks, so local HBM/L2 work overlaps the
// NVLink round trips.  Measured (profiles/r02_p2p_notes.md): visiting ALL homes round-robin by chunk is fine with one
// peer (N=2: 0.78 vs 0.94 ms) but collapses with three or seven (N=4: 4.99 vs 1.78 ms, N=8: 10 ms) - whatever the
// cause (peer-aperture translation is the suspect), a rank must stream from one peer at a time.
// pref[k] = first chunk of source slot k in plain sequential order (slot W-1 = local), m = interleaved local chunks.
__device__ __forceinline__ void chunk_of(int c, int W, int m, const int *pref, bool round_robin, int &k, int &off) {
    if (round_robin) {                                   // measurement mode: every source by turns while all have chunks
        if (c < m * W) { k = c % W; off = (c / W) << 5; return; }
        int c2 = c - m * W, kk = 0;                      // leftovers source by source
        for (;; ++kk) {
            const int left = pref[kk + 1] - pref[kk] - m;
            if (c2 < left) break;
            c2 -= left;
        }
        k = kk; off = (m + c2) << 5;
        return;
    }
    const int R = pref[W - 1];                           // remote chunks; local chunks: pref[W] - R
    int r;                                               // linear index into the remote stream, or -1 = local chunk q
    int q = 0;
    if (W == 1) { k = 0; off = c << 5; return; }
    if (c < W * m) {
        const int qq = c / W, pos = c % W;
        if (pos == W - 1) { r = -1; q = qq; } else r = qq * (W - 1) + pos;
    } else {
        const int c2 = c - W * m, remR = R - m * (W - 1);
        if (c2 < remR) r = m * (W - 1) + c2; else { r = -1; q = m + (c2 - remR); }
    }
    if (r < 0) { k = W - 1; off = q << 5; return; }
    k = 0;
    while (r >= pref[k + 1]) ++k;
    off = (r - pref[k]) << 5;
}

template <int CPL>
struct P2PRows {
    float4 u[CPL], i[CPL], j[CPL];
    int tu, ti, tj;
};

struct P2PParams {
    b200rec_p2p_step_args a;
    unsigned int *work;
};

template <int CPL, bool UNIQ, bool LOSS>
__global__ void __launch_bounds__(256, CPL == 1 ? 3 : (CPL == 2 ? 2 : 1)) p2p_step_kernel(const __grid_constant__ P2PParams p) {
    __shared__ int s_cnt[B200REC_MAX_RANKS];        // triples from source order[k]
    __shared__ int s_pref[B200REC_MAX_RANKS + 1];   // chunk prefix over sources in visiting order
    __shared__ int s_src[B200REC_MAX_RANKS];
    __shared__ float *s_U[B200REC_MAX_RANKS];
    __shared__ float *s_V[B200REC_MAX_RANKS];
    __shared__ const int32_t *s_iu[B200REC_MAX_RANKS], *s_ii[B200REC_MAX_RANKS], *s_ij[B200REC_MAX_RANKS];
    __shared__ int s_bounds[B200REC_MAX_RANKS + 1];
    __shared__ int s_rr;
    const b200rec_p2p_step_args &a = p.a;
    const int W = a.world, me = a.rank;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < W) {
        // visiting order: start with the next rank so that at any moment different ranks pull from different homes
        const int k = threadIdx.x, s = (me + 1 + k) % W;
        s_src[k] = s;
        s_cnt[k] = *a.in_cnt[s];                     // peer load (local for s == me)
        s_iu[k] = a.in_u[s]; s_ii[k] = a.in_i[s]; s_ij[k] = a.in_j[s];
    }
    if (threadIdx.x < B200REC_MAX_RANKS) { s_U[threadIdx.x] = a.U_peer[threadIdx.x]; s_V[threadIdx.x] = a.V_peer[threadIdx.x]; }
    if (threadIdx.x <= W) s_bounds[threadIdx.x] = a.item_bounds[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        // s_pref: chunk prefix over the sources in visiting order (slots 0..W-2 are the remote homes me+1, me+2, ...,
        // slot W-1 is this rank itself); s_rr: how many local chunks are interleaved with the remote stream (chunk_of)
        int acc = 0;
        for (int k = 0; k < W; ++k) { s_pref[k] = acc; acc += (s_cnt[k] + 31) >> 5; }
        s_pref[W] = acc;
        const int R = s_pref[W - 1], nl = acc - R;
        int m = (W > 1) ? min(nl, R / (W - 1)) : 0;
        if (a.flags & B200REC_F_P2P_PURE_SEQUENTIAL) m = 0;
        if (a.flags & B200REC_F_P2P_ROUND_ROBIN) {
            m = 0x7fffffff;
            for (int k = 0; k < W; ++k) m = min(m, (s_cnt[k] + 31) >> 5);
        }
        s_rr = m;
        if (blockIdx.x == 0 && a.n_processed) {
            int tot = 0;
            for (int k = 0; k < W; ++k) tot += s_cnt[k];
            *a.n_processed = tot;
        }
    }
    __syncthreads();
    const int n_chunks = s_pref[W];
    const int ld = a.ld, d4 = a.ld >> 2;
    const int item_lo = s_bounds[me], item_hi = s_bounds[me + 1];
    float *const Vloc = s_V[me];
    const int head = a.head;
    float *const Vh = a.Vh, *const dVh = a.dVh;
    const float c_g = a.lr * a.inv_batch;                 // delta = c_g*(1-s) * other + c_r * self
    const float c_r = -a.lr * a.reg * a.inv_batch;
    float loss_local = 0.f;

    unsigned int *const work = p.work;
    int c = 0, c_next = 0;
    if (lane == 0) c = (int)atomicAdd(work, 1u);
    c = __shfl_sync(0xffffffffu, c, 0);
    for (; c < n_chunks; c = c_next) {
        if (lane == 0) c_next = (int)atomicAdd(work, 1u);   // latency hides behind this chunk's rows
        int k, off;
        chunk_of(c, W, s_rr, s_pref, (a.flags & B200REC_F_P2P_ROUND_ROBIN) != 0, k, off);
        const int n_here = min(32, s_cnt[k] - off);
        int u = 0, i = 0, j = 0;
        if (lane < n_here) { u = s_iu[k][off + lane]; i = s_ii[k][off + lane]; j = s_ij[k][off + lane]; }
        float *const Uhome = s_U[s_src[k]];

        P2PRows<CPL> r0, r1, r2;
        // where an item row is READ (head replica / own shard / a peer's shard) and where its update GOES (the head's
        // delta buffer, or the row itself)
        auto vptr = [&](int id, bool write) -> float * {
            if (id < head) return (write ? dVh : Vh) + (int64_t)id * ld;
            if (id >= item_lo && id < item_hi) return Vloc + (int64_t)(id - item_lo) * ld;
            int o = 0;
            for (int q = 1; q < W; ++q) o += (id >= s_bounds[q]);
            return s_V[o] + (int64_t)(id - s_bounds[o]) * ld;
        };
        auto load = [&](P2PRows<CPL> &r, int it) {
            if (it < n_here) {
                r.tu = __shfl_sync(0xffffffffu, u, it);
                r.ti = __shfl_sync(0xffffffffu, i, it);
                r.tj = __shfl_sync(0xffffffffu, j, it);
                const float *pu = Uhome + (int64_t)r.tu * ld + lane * 4;
                const float *pi = vptr(r.ti, false) + lane * 4;
                const float *pj = vptr(r.tj, false) + lane * 4;
#pragma unroll
                for (int q = 0; q < CPL; ++q) {
                    if (lane + 32 * q < d4) { r.u[q] = ld4(pu + 128 * q); r.i[q] = ld4(pi + 128 * q); r.j[q] = ld4(pj + 128 * q); }
                    else r.u[q] = r.i[q] = r.j[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        };
        auto compute = [&](const P2PRows<CPL> &r, int it) {
            if (it < n_here) {
                float4 df[CPL];
                float part = 0.f;
#pragma unroll
                for (int q = 0; q < CPL; ++q) {
                    df[q] = make_float4(r.i[q].x - r.j[q].x, r.i[q].y - r.j[q].y, r.i[q].z - r.j[q].z, r.i[q].w - r.j[q].w);
                    part = fmaf(r.u[q].x, df[q].x, part); part = fmaf(r.u[q].y, df[q].y, part);
                    part = fmaf(r.u[q].z, df[q].z, part); part = fmaf(r.u[q].w, df[q].w, part);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                const float x = part;
                const float s = __frcp_rn(1.f + __expf(-x));       // sigmoid(x); 1 - s saturates like the reference's fp32
                const float a1 = c_g * (1.f - s);                   // = -lr * g
                if (LOSS) loss_local += (x < -15.f) ? -x : -__logf(s);
                float *pu = Uhome + (int64_t)r.tu * ld + lane * 4;
                float *pi = vptr(r.ti, true) + lane * 4;
                float *pj = vptr(r.tj, true) + lane * 4;
#pragma unroll
                for (int q = 0; q < CPL; ++q) {
                    if (lane + 32 * q < d4) {
                        float4 du, di, dj;
                        du.x = fmaf(a1, df[q].x, c_r * r.u[q].x); du.y = fmaf(a1, df[q].y, c_r * r.u[q].y);
                        du.z = fmaf(a1, df[q].z, c_r * r.u[q].z); du.w = fmaf(a1, df[q].w, c_r * r.u[q].w);
                        di.x = fmaf(a1, r.u[q].x, c_r * r.i[q].x); di.y = fmaf(a1, r.u[q].y, c_r * r.i[q].y);
                        di.z = fmaf(a1, r.u[q].z, c_r * r.i[q].z); di.w = fmaf(a1, r.u[q].w, c_r * r.i[q].w);
                        dj.x = fmaf(-a1, r.u[q].x, c_r * r.j[q].x); dj.y = fmaf(-a1, r.u[q].y, c_r * r.j[q].y);
                        dj.z = fmaf(-a1, r.u[q].z, c_r * r.j[q].z); dj.w = fmaf(-a1, r.u[q].w, c_r * r.j[q].w);
                        if (UNIQ) st4(pu + 128 * q, make_float4(r.u[q].x + du.x, r.u[q].y + du.y, r.u[q].z + du.z, r.u[q].w + du.w));
                        else red4(pu + 128 * q, du);
                        red4(pi + 128 * q, di);
                        red4(pj + 128 * q, dj);
                    }
                }
            }
        };
        load(r0, 0);
        load(r1, 1);
        for (int it = 0; it < n_here; it += 3) {
            load(r2, it + 2);
            compute(r0, it);
            load(r0, it + 3);
            compute(r1, it + 1);
            load(r1, it + 4);
            compute(r2, it + 2);
        }
        c_next = __shfl_sync(0xffffffffu, c_next, 0);
    }
    if (LOSS) {
        if (lane == 0 && loss_local != 0.f) atomicAdd(a.loss_sum, (double)loss_local);
    }
    __threadfence_system();   // peer stores of this thread are performed before the kernel (and the next barrier) ends
}

// Same step with a row held by a sub-warp group of G lanes (ld = 128 floats: 32/G float4 per lane), 32/G triples per warp
// at a time: more independent loads in flight per lane (what hides the ~2 us NVLink round trip of a remote user row),
// a 3-stage instead of a 5-stage reduction, and the scalar part paid once per 32/G triples (cf. bpr_step_group_kernel).
template <int G, bool UNIQ, bool LOSS>
__global__ void __launch_bounds__(256, 3) p2p_step_group_kernel(const __grid_constant__ P2PParams p) {
    constexpr int F4 = 32, CPL = F4 / G, TPW = 32 / G, ITERS = 32 / TPW, LD = F4 * 4;
    const bool no_uwrite = (p.a.flags & B200REC_F_P2P_NO_UWRITE) != 0, no_uread = (p.a.flags & B200REC_F_P2P_NO_UREAD) != 0;
    __shared__ int s_cnt[B200REC_MAX_RANKS];
    __shared__ int s_pref[B200REC_MAX_RANKS + 1];
    __shared__ int s_src[B200REC_MAX_RANKS];
    __shared__ float *s_U[B200REC_MAX_RANKS];
    __shared__ float *s_V[B200REC_MAX_RANKS];
    __shared__ const int32_t *s_iu[B200REC_MAX_RANKS], *s_ii[B200REC_MAX_RANKS], *s_ij[B200REC_MAX_RANKS];
    __shared__ int s_bounds[B200REC_MAX_RANKS + 1];
    __shared__ int s_rr;
    const b200rec_p2p_step_args &a = p.a;
    const int W = a.world, me = a.rank;
    const int lane = threadIdx.x & 31, sl = lane % G, sg = lane / G;
    if (threadIdx.x < W) {
        const int k = threadIdx.x, s = (me + 1 + k) % W;
        s_src[k] = s;
        s_cnt[k] = *a.in_cnt[s];
        s_iu[k] = a.in_u[s]; s_ii[k] = a.in_i[s]; s_ij[k] = a.in_j[s];
    }
    if (threadIdx.x < B200REC_MAX_RANKS) { s_U[threadIdx.x] = a.U_peer[threadIdx.x]; s_V[threadIdx.x] = a.V_peer[threadIdx.x]; }
    if (threadIdx.x <= W) s_bounds[threadIdx.x] = a.item_bounds[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        // s_pref: chunk prefix over the sources in visiting order (slots 0..W-2 are the remote homes me+1, me+2, ...,
        // slot W-1 is this rank itself); s_rr: how many local chunks are interleaved with the remote stream (chunk_of)
        int acc = 0;
        for (int k = 0; k < W; ++k) { s_pref[k] = acc; acc += (s_cnt[k] + 31) >> 5; }
        s_pref[W] = acc;
        const int R = s_pref[W - 1], nl = acc - R;
        int m = (W > 1) ? min(nl, R / (W - 1)) : 0;
        if (a.flags & B200REC_F_P2P_PURE_SEQUENTIAL) m = 0;
        if (a.flags & B200REC_F_P2P_ROUND_ROBIN) {
            m = 0x7fffffff;
            for (int k = 0; k < W; ++k) m = min(m, (s_cnt[k] + 31) >> 5);
        }
        s_rr = m;
        if (blockIdx.x == 0 && a.n_processed) {
            int tot = 0;
            for (int k = 0; k < W; ++k) tot += s_cnt[k];
            *a.n_processed = tot;
        }
    }
    __syncthreads();
    const int n_chunks = s_pref[W];
    const int item_lo = s_bounds[me], item_hi = s_bounds[me + 1];
    float *const Vloc = s_V[me];
    const int head = a.head;
    float *const Vh = a.Vh, *const dVh = a.dVh;
    const float c_g = a.lr * a.inv_batch;
    const float c_r = -a.lr * a.reg * a.inv_batch;
    float loss_local = 0.f;
    unsigned int *const work = p.work;
    int c = 0, c_next = 0;
    if (lane == 0) c = (int)atomicAdd(work, 1u);
    c = __shfl_sync(0xffffffffu, c, 0);
    for (; c < n_chunks; c = c_next) {
        if (lane == 0) c_next = (int)atomicAdd(work, 1u);
        int k, off;
        chunk_of(c, W, s_rr, s_pref, (a.flags & B200REC_F_P2P_ROUND_ROBIN) != 0, k, off);
        const int n_here = min(32, s_cnt[k] - off);
        int u = 0, i = 0, j = 0;
        if (lane < n_here) { u = s_iu[k][off + lane]; i = s_ii[k][off + lane]; j = s_ij[k][off + lane]; }
        float *const Uhome = s_U[s_src[k]];
        auto vptr = [&](int id, bool write) -> float * {
            if (id < head) return (write ? dVh : Vh) + (int64_t)id * LD;
            if (id >= item_lo && id < item_hi) return Vloc + (int64_t)(id - item_lo) * LD;
            int o = 0;
            for (int q = 1; q < W; ++q) o += (id >= s_bounds[q]);
            return s_V[o] + (int64_t)(id - s_bounds[o]) * LD;
        };
        // Hot positive of the chunk.  A rank owns the item rows of its range, so under a Zipf catalogue the rank holding
        // THE most popular item sends a third of its positive-row reductions to one row (4 L2 lines) - they serialise
        // in the slice's atomic unit: measured 0.99 ms per 1M triples on that rank against 0.46 ms elsewhere (N=4), the
        // whole step waits for it.  The triples of the chunk that share the most frequent positive are therefore
        // processed first, their item update is summed in registers and leaves as ONE reduction per chunk, and the row
        // is read once.
        const unsigned vmask = __ballot_sync(0xffffffffu, lane < n_here);
        const unsigned same = __match_any_sync(0xffffffffu, lane < n_here ? i : -1 - lane);
        const int cnt_same = (lane < n_here) ? __popc(same) : 0;
        const int mx = __reduce_max_sync(0xffffffffu, cnt_same);
        unsigned hot = 0;
        if (mx >= 3) hot = __shfl_sync(0xffffffffu, same, __ffs(__ballot_sync(0xffffffffu, cnt_same == mx)) - 1);
        const int nh = __popc(hot);
        const unsigned cold = vmask & ~hot;
        const int i_hot = nh ? __shfl_sync(0xffffffffu, i, __ffs(hot) - 1) : -1;
        float4 hrow[CPL], hacc[CPL];
        if (nh) {
            const float *ph = vptr(i_hot, false) + sl * 4;
#pragma unroll
            for (int q = 0; q < CPL; ++q) { hrow[q] = ld4(ph + q * G * 4); hacc[q] = make_float4(0.f, 0.f, 0.f, 0.f); }
        }
        // the user row (the one that may come over NVLink: ~2.5 us) is fetched ONE ITERATION AHEAD, so its round trip
        // overlaps the item-row work of the current pair of triples
        auto src_of = [&](int pos) -> int {                      // lane that holds the triple processed at `pos`
            if (pos >= n_here) return 0;
            return pos < nh ? __fns(hot, 0, pos + 1) : __fns(cold, 0, pos - nh + 1);
        };
        float4 nu[CPL];
        int src_nx = src_of(sg);
        int tu_nx = __shfl_sync(0xffffffffu, u, src_nx);
        if (sg < n_here && !no_uread) {
#pragma unroll
            for (int q = 0; q < CPL; ++q) nu[q] = ld4(Uhome + (int64_t)tu_nx * LD + sl * 4 + q * G * 4);
        }
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            if (it * TPW >= n_here) break;                       // warp-uniform
            const int pos = it * TPW + sg;                       // position in the processing order: hot triples first
            const bool ok = pos < n_here;
            const bool is_hot = pos < nh;
            const int src = src_nx;
            const int tu = tu_nx;
            const int ti = __shfl_sync(0xffffffffu, i, src);
            const int tj = __shfl_sync(0xffffffffu, j, src);
            float4 ru[CPL], ri[CPL], rj[CPL];
#pragma unroll
            for (int q = 0; q < CPL; ++q) ru[q] = nu[q];
            src_nx = src_of(pos + TPW);
            tu_nx = __shfl_sync(0xffffffffu, u, src_nx);
            if (pos + TPW < n_here && !no_uread) {
#pragma unroll
                for (int q = 0; q < CPL; ++q) nu[q] = ld4(Uhome + (int64_t)tu_nx * LD + sl * 4 + q * G * 4);
            }
            float part = 0.f;
            if (ok) {
                const float *pj = vptr(tj, false) + sl * 4;
                if (no_uread) {
#pragma unroll
                    for (int q = 0; q < CPL; ++q) ru[q] = make_float4(0.01f, 0.01f, 0.01f, 0.01f);
                }
                if (is_hot) {
#pragma unroll
                    for (int q = 0; q < CPL; ++q) ri[q] = hrow[q];
                } else {
                    const float *pi = vptr(ti, false) + sl * 4;
#pragma unroll
                    for (int q = 0; q < CPL; ++q) ri[q] = ld4(pi + q * G * 4);
                }
#pragma unroll
                for (int q = 0; q < CPL; ++q) rj[q] = ld4(pj + q * G * 4);
#pragma unroll
                for (int q = 0; q < CPL; ++q) {
                    part = fmaf(ru[q].x, ri[q].x - rj[q].x, part); part = fmaf(ru[q].y, ri[q].y - rj[q].y, part);
                    part = fmaf(ru[q].z, ri[q].z - rj[q].z, part); part = fmaf(ru[q].w, ri[q].w - rj[q].w, part);
                }
            }
#pragma unroll
            for (int o = G >> 1; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (ok) {
                const float x = part;
                const float s = __frcp_rn(1.f + __expf(-x));
                const float a1 = c_g * (1.f - s);
                if (LOSS && sl == 0) loss_local += (x < -15.f) ? -x : -__logf(s);
                float *pu = Uhome + (int64_t)tu * LD + sl * 4;
                float *pi = is_hot ? nullptr : vptr(ti, true) + sl * 4;
                float *pj = vptr(tj, true) + sl * 4;
#pragma unroll
                for (int q = 0; q < CPL; ++q) {
                    const float4 vu = ru[q], vi = ri[q], vj = rj[q];
                    float4 du, di, dj;
                    du.x = fmaf(a1, vi.x - vj.x, c_r * vu.x); du.y = fmaf(a1, vi.y - vj.y, c_r * vu.y);
                    du.z = fmaf(a1, vi.z - vj.z, c_r * vu.z); du.w = fmaf(a1, vi.w - vj.w, c_r * vu.w);
                    di.x = fmaf(a1, vu.x, c_r * vi.x); di.y = fmaf(a1, vu.y, c_r * vi.y);
                    di.z = fmaf(a1, vu.z, c_r * vi.z); di.w = fmaf(a1, vu.w, c_r * vi.w);
                    dj.x = fmaf(-a1, vu.x, c_r * vj.x); dj.y = fmaf(-a1, vu.y, c_r * vj.y);
                    dj.z = fmaf(-a1, vu.z, c_r * vj.z); dj.w = fmaf(-a1, vu.w, c_r * vj.w);
                    if (!no_uwrite) {
                        if (UNIQ) st4(pu + q * G * 4, make_float4(vu.x + du.x, vu.y + du.y, vu.z + du.z, vu.w + du.w));
                        else red4(pu + q * G * 4, du);
                    }
                    if (is_hot) { hacc[q].x += di.x; hacc[q].y += di.y; hacc[q].z += di.z; hacc[q].w += di.w; }
                    else red4(pi + q * G * 4, di);
                    red4(pj + q * G * 4, dj);
                }
            }
        }
        if (nh) {   // one reduction for all hot triples of the chunk: fold the groups' partial sums, group 0 writes
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
#pragma unroll
                for (int o = G; o < 32; o <<= 1) {
                    hacc[q].x += __shfl_xor_sync(0xffffffffu, hacc[q].x, o); hacc[q].y += __shfl_xor_sync(0xffffffffu, hacc[q].y, o);
                    hacc[q].z += __shfl_xor_sync(0xffffffffu, hacc[q].z, o); hacc[q].w += __shfl_xor_sync(0xffffffffu, hacc[q].w, o);
                }
            }
            if (sg == 0) {
                float *ph = vptr(i_hot, true) + sl * 4;
#pragma unroll
                for (int q = 0; q < CPL; ++q) red4(ph + q * G * 4, hacc[q]);
            }
        }
        c_next = __shfl_sync(0xffffffffu, c_next, 0);
    }
    if (LOSS) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) loss_local += __shfl_xor_sync(0xffffffffu, loss_local, o);
        if (lane == 0 && loss_local != 0.f) atomicAdd(a.loss_sum, (double)loss_local);
    }
    __threadfence_system();
}

static int next_work_counter(cudaStream_t s, unsigned int **out) {
    static unsigned int *ring[64] = {nullptr};
    static unsigned int slot[64] = {0};
    int dev = 0;
    B200_CUDA(cudaGetDevice(&dev));
    B200_REQUIRE(dev >= 0 && dev < 64, B200REC_EUNSUPPORTED, "p2p: device index %d out of range", dev);
    if (!ring[dev]) B200_CUDA(cudaMalloc(&ring[dev], 256 * sizeof(unsigned int)));
    *out = ring[dev] + (slot[dev]++ & 255u);
    B200_CUDA(cudaMemsetAsync(*out, 0, sizeof(unsigned int), s));
    return B200REC_OK;
}

}  // namespace b200

using namespace b200;

extern "C" int b200rec_p2p_route(const b200rec_p2p_route_args *args, void *stream) {
    B200_REQUIRE(args != nullptr, B200REC_EINVAL, "p2p_route: args is NULL");
    const b200rec_p2p_route_args &a = *args;
    B200_REQUIRE(a.world >= 1 && a.world <= B200REC_MAX_RANKS && a.rank >= 0 && a.rank < a.world, B200REC_EINVAL,
                 "p2p_route: need 1 <= world <= %d and 0 <= rank < world", B200REC_MAX_RANKS);
    B200_REQUIRE(a.out_u && a.out_i && a.out_j && a.out_cnt && a.cap >= a.B && a.B >= 0, B200REC_EINVAL,
                 "p2p_route: outbox missing or smaller than the batch (cap=%d, B=%d)", a.cap, a.B);
    B200_REQUIRE(a.B == 0 || a.users, B200REC_EINVAL, "p2p_route: users is NULL");
    B200_REQUIRE((a.pos && a.neg) || (a.csr_indptr && a.csr_indices), B200REC_EINVAL,
                 "p2p_route: on-device sampling needs the CSR shard");
    B200_REQUIRE(a.head >= 0 && a.item_bounds[0] == a.head, B200REC_EINVAL,
                 "p2p_route: item_bounds must start at head (%d), got %d", a.head, a.item_bounds[0]);
    for (int r = 0; r < a.world; ++r)
        B200_REQUIRE(a.item_bounds[r] <= a.item_bounds[r + 1], B200REC_EINVAL, "p2p_route: item_bounds must be non-decreasing");
    for (int r = 0; r < a.world && !a.neg; ++r)
        B200_REQUIRE(a.head > 0 || a.item_bounds[r] < a.item_bounds[r + 1], B200REC_EINVAL,
                     "p2p_route: an empty item shard cannot supply negatives");
    cudaStream_t s = (cudaStream_t)stream;
    B200_CUDA(cudaMemsetAsync(a.out_cnt, 0, sizeof(int32_t) * a.world, s));
    if (a.B == 0) return B200REC_OK;
    const int64_t tile = (int64_t)kRouteThreads * kRouteTilesPerCta;
    const int64_t need = (a.B + tile - 1) / tile, cap = (int64_t)sm_count() * 8;
    p2p_route_kernel<<<(int)(need < cap ? need : cap), kRouteThreads, 0, s>>>(a);
    B200_LAUNCH_CHECK();
    return B200REC_OK;
}

extern "C" int b200rec_p2p_step(const b200rec_p2p_step_args *args, void *stream) {
    B200_REQUIRE(args != nullptr, B200REC_EINVAL, "p2p_step: args is NULL");
    const b200rec_p2p_step_args &a = *args;
    B200_REQUIRE(a.world >= 1 && a.world <= B200REC_MAX_RANKS && a.rank >= 0 && a.rank < a.world, B200REC_EINVAL,
                 "p2p_step: need 1 <= world <= %d and 0 <= rank < world", B200REC_MAX_RANKS);
    B200_REQUIRE(a.d >= 1 && a.ld >= a.d && (a.ld % 4) == 0 && a.ld <= 512, B200REC_EINVAL,
                 "p2p_step: need 1 <= d <= ld <= 512, ld %% 4 == 0 (d=%d ld=%d)", a.d, a.ld);
    B200_REQUIRE(a.inv_batch > 0.f, B200REC_EINVAL, "p2p_step: inv_batch (1 / global batch) is required");
    B200_REQUIRE(a.head >= 0 && a.item_bounds[0] == a.head && (a.head == 0 || (a.Vh && a.dVh)), B200REC_EINVAL,
                 "p2p_step: head rows need Vh and dVh, and item_bounds[0] == head");
    for (int r = 0; r < a.world; ++r) {
        B200_REQUIRE(a.U_peer[r] && a.V_peer[r] && a.in_u[r] && a.in_i[r] && a.in_j[r] && a.in_cnt[r], B200REC_EINVAL,
                     "p2p_step: peer table entry %d is NULL", r);
        B200_REQUIRE(((uintptr_t)a.U_peer[r] % 16) == 0 && ((uintptr_t)a.V_peer[r] % 16) == 0, B200REC_EINVAL,
                     "p2p_step: tables must be 16-byte aligned");
        B200_REQUIRE(a.item_bounds[r] <= a.item_bounds[r + 1], B200REC_EINVAL, "p2p_step: bad item_bounds");
    }
    cudaStream_t s = (cudaStream_t)stream;
    P2PParams p;
    p.a = a;
    int rc = next_work_counter(s, &p.work);
    if (rc) return rc;
    const int cpl = (a.ld / 4 + 31) / 32;
    const bool uniq = (a.flags & B200REC_F_USERS_UNIQUE) != 0, loss = a.loss_sum != nullptr;
    {   // ld == 128: group kernel (B200REC_P2P_VARIANT=0 selects the warp-per-row kernel)
        const char *var = getenv("B200REC_P2P_VARIANT");
        const int vg = var ? atoi(var) : 16;
        if (a.ld == 128 && vg > 0) {
            auto launch = [&](auto kern) -> int {
                int occ = 0;
                B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0));
                if (occ < 1) occ = 1;
                kern<<<sm_count() * occ, 256, 0, s>>>(p);
                return B200REC_OK;
            };
            int rc_g;
#define B200_P2PG(GG)                                                                                            \
    if (uniq) { rc_g = loss ? launch(p2p_step_group_kernel<GG, true, true>) : launch(p2p_step_group_kernel<GG, true, false>); } \
    else { rc_g = loss ? launch(p2p_step_group_kernel<GG, false, true>) : launch(p2p_step_group_kernel<GG, false, false>); }
            if (vg == 16) { B200_P2PG(16) } else { B200_P2PG(8) }
#undef B200_P2PG
            if (rc_g) return rc_g;
            B200_LAUNCH_CHECK();
            return B200REC_OK;
        }
    }
    const int grid = sm_count() * (cpl == 1 ? 3 : (cpl == 2 ? 2 : 1));
#define B200_P2P(C)                                                                     \
    {                                                                                   \
        if (uniq) { if (loss) p2p_step_kernel<C, true, true><<<grid, 256, 0, s>>>(p);   \
                    else p2p_step_kernel<C, true, false><<<grid, 256, 0, s>>>(p); }     \
        else { if (loss) p2p_step_kernel<C, false, true><<<grid, 256, 0, s>>>(p);       \
               else p2p_step_kernel<C, false, false><<<grid, 256, 0, s>>>(p); }         \
    }
    switch (cpl) {
        case 1: B200_P2P(1) break;
        case 2: B200_P2P(2) break;
        case 3: B200_P2P(3) break;
        default: B200_P2P(4) break;
    }
#undef B200_P2P
    B200_LAUNCH_CHECK();
    return B200REC_OK;
}

// ---- peer memory plumbing (CUDA IPC): one exportable allocation per rank, opened by every other rank -----------
extern "C" int b200rec_peer_alloc(int64_t bytes, void **dev_ptr) {
    B200_REQUIRE(dev_ptr && bytes > 0, B200REC_EINVAL, "peer_alloc: bad argument");
    cudaError_t e = cudaMalloc(dev_ptr, (size_t)bytes);
    if (e != cudaSuccess) { set_error("peer_alloc: cudaMalloc(%lld) -> %s", (long long)bytes, cudaGetErrorString(e)); return B200REC_ENOMEM; }
    return B200REC_OK;
}
extern "C" int b200rec_peer_free(void *dev_ptr) {
    if (dev_ptr) B200_CUDA(cudaFree(dev_ptr));
    return B200REC_OK;
}
extern "C" int b200rec_peer_export(void *dev_ptr, void *handle64) {
    B200_REQUIRE(dev_ptr && handle64, B200REC_EINVAL, "peer_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == B200REC_PEER_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    B200_CUDA(cudaIpcGetMemHandle(&h, dev_ptr));
    memcpy(handle64, &h, sizeof(h));
    return B200REC_OK;
}
extern "C" int b200rec_peer_import(const void *handle64, void **dev_ptr) {
    B200_REQUIRE(dev_ptr && handle64, B200REC_EINVAL, "peer_import: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    B200_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return B200REC_OK;
}
extern "C" int b200rec_peer_close(void *dev_ptr) {
    if (dev_ptr) B200_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return B200REC_OK;
}
extern "C" int b200rec_peer_copy(void *dst, const void *src, int64_t bytes, void *stream) {
    B200_REQUIRE(dst && src && bytes >= 0, B200REC_EINVAL, "peer_copy: bad argument");
    if (bytes == 0) return B200REC_OK;
    B200_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return B200REC_OK;
}
