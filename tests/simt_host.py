"""Test infrastructure: a minimal SIMT emulator that runs the fused BPR step kernels of csrc/bpr_step.cu ON THE HOST,
compiled from their own source text, so `-m "not gpu"` tests can execute the kernels' real code (warp-cooperative id
exchange, sub-warp row groups, chunk handling, every sink, the on-device sampler) against the numpy oracle on a box
without a GPU.  Nothing here is shipped or imported by the product; the GPU parity suite stays the authority for the
hardware behaviour (memory model, tcgen05 / TMA paths are NOT emulated).

How: every lane of a warp is a host thread; the warp intrinsics the kernels use (`__shfl_sync`, `__shfl_xor_sync`,
`__ballot_sync`, `__syncwarp`) are rendezvous over a 32-party barrier with an exchange array; `red4` is four CAS float
adds, `atomicAdd` the gcc builtin / a CAS loop; `__expf`, `__frcp_rn`, `__logf` map to libm (a few ulp from the device's
fast paths - far inside the 2e-5 parity tolerance).  The 8 warps of a CTA run concurrently, CTAs one after another.
Taken verbatim from the sources: mix64, rng_u32 (common.cuh); sampler.cuh; BprParams, bpr_grad, softplus_neg, group_sum,
sink_chunk, bpr_step_ldg_kernel, RowSet, bpr_step_fast_kernel, GroupSet, bpr_step_group_kernel, bpr_apply_kernel
(bpr_step.cu).  The launch parameters (chunk size, chunk count, 1/B, work counter) are set the way b200rec_bpr_step does.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "recsys_pytorch_b200", "csrc")

_PRELUDE = r'''
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>
#include <thread>
#include <vector>
#include <map>
#include <mutex>
#include "b200rec.h"

struct float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 v = {x, y, z, w}; return v; }
struct Dim3 { int x, y; };
static thread_local Dim3 blockIdx, blockDim, threadIdx, gridDim;

struct WarpCtx {
    pthread_barrier_t bar; uint64_t xchg[32];
    std::mutex mu; std::map<unsigned, pthread_barrier_t *> sub;          // rendezvous of a sub-warp group (partial member mask)
    pthread_barrier_t *group(unsigned mask) {
        std::lock_guard<std::mutex> g(mu);
        auto it = sub.find(mask);
        if (it != sub.end()) return it->second;
        pthread_barrier_t *b = new pthread_barrier_t;
        pthread_barrier_init(b, nullptr, __builtin_popcount(mask));
        sub[mask] = b; return b;
    }
    ~WarpCtx() { for (auto &kv : sub) { pthread_barrier_destroy(kv.second); delete kv.second; } }
};
struct BlockCtx { pthread_barrier_t bar; };
static thread_local WarpCtx *t_warp;
static thread_local BlockCtx *t_block;
static thread_local int t_lane;
static inline void warp_bar() { pthread_barrier_wait(&t_warp->bar); }
static inline void __syncthreads() { pthread_barrier_wait(&t_block->bar); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

// member mask = the whole warp, or one sub-warp group whose lanes (and only they) all make the call (csrc/spmm.cu)
template <class T> static inline T __shfl_sync(unsigned mask, T v, int src) {
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    pthread_barrier_t *b = (mask == 0xffffffffu) ? &t_warp->bar : ((mask & (mask - 1)) ? t_warp->group(mask) : nullptr);
    t_warp->xchg[t_lane] = raw; if (b) pthread_barrier_wait(b);
    const uint64_t r = t_warp->xchg[src & 31]; if (b) pthread_barrier_wait(b);
    T out; memcpy(&out, &r, sizeof(T)); return out;
}
template <class T> static inline T __shfl_xor_sync(unsigned m, T v, int o) { return __shfl_sync(m, v, t_lane ^ o); }
static inline unsigned __ballot_sync(unsigned, int pred) {
    t_warp->xchg[t_lane] = pred ? 1 : 0; warp_bar();
    unsigned r = 0; for (int l = 0; l < 32; ++l) r |= (unsigned)t_warp->xchg[l] << l;
    warp_bar(); return r;
}
static inline void __syncwarp() { warp_bar(); }
// lanes holding the same value (the kernels only pass full masks)
static inline unsigned __match_any_sync(unsigned, int v) {
    t_warp->xchg[t_lane] = (uint64_t)(uint32_t)v; warp_bar();
    unsigned r = 0; for (int l = 0; l < 32; ++l) r |= (unsigned)(t_warp->xchg[l] == (uint64_t)(uint32_t)v) << l;
    warp_bar(); return r;
}
static inline int __reduce_max_sync(unsigned, int v) {
    t_warp->xchg[t_lane] = (uint64_t)(int64_t)v; warp_bar();
    int r = v; for (int l = 0; l < 32; ++l) { const int o = (int)(int64_t)t_warp->xchg[l]; r = o > r ? o : r; }
    warp_bar(); return r;
}
static inline int __reduce_add_sync(unsigned, int v) {
    t_warp->xchg[t_lane] = (uint64_t)(int64_t)v; warp_bar();
    int r = 0; for (int l = 0; l < 32; ++l) r += (int)(int64_t)t_warp->xchg[l];
    warp_bar(); return r;
}
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
// position of the offset-th set bit of mask at or above `base` (offset > 0), 0xffffffff when there is none (CUDA __fns)
static inline unsigned __fns(unsigned mask, unsigned base, int offset) {
    if (offset <= 0) { fprintf(stderr, "simt_host: __fns only emulated for offset > 0\n"); abort(); }
    for (unsigned b = base; b < 32; ++b) if ((mask >> b) & 1u) { if (--offset == 0) return b; }
    return 0xffffffffu;
}
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }

static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline int atomicExch(int *p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
static inline void atomic_addf(float *p, float v);
static inline float atomicAdd(float *p, float v) { atomic_addf(p, v); return 0.f; }
static inline void atomic_addf(float *p, float v) {
    uint32_t *u = reinterpret_cast<uint32_t *>(p), old = __atomic_load_n(u, __ATOMIC_RELAXED), neu;
    do { float f; memcpy(&f, &old, 4); f += v; memcpy(&neu, &f, 4); }
    while (!__atomic_compare_exchange_n(u, &old, neu, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
}
static inline double atomicAdd(double *p, double v) {
    uint64_t *u = reinterpret_cast<uint64_t *>(p), old = __atomic_load_n(u, __ATOMIC_RELAXED), neu; double f;
    do { memcpy(&f, &old, 8); double g = f + v; memcpy(&neu, &g, 8); }
    while (!__atomic_compare_exchange_n(u, &old, neu, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    return f;
}
static inline float emu_expf(float x) { return expf(x); }      // the kernels' __expf / __logf / __frcp_rn (glibc owns
static inline float emu_logf(float x) { return logf(x); }      // the first two names, so the source text is renamed)
static inline float emu_frcp_rn(float x) { return 1.0f / x; }

struct __nv_bfloat16 { uint16_t v; };
#define B200REC_FAST_MINB 3
#define ABL(bit) 0

namespace b200 {
static inline float4 ld4(const float *p) { float4 v; memcpy(&v, p, 16); return v; }
static inline void st4(float *p, float4 v) { memcpy(p, &v, 16); }
static inline void red4(float *p, float4 v) { atomic_addf(p, v.x); atomic_addf(p + 1, v.y); atomic_addf(p + 2, v.z); atomic_addf(p + 3, v.w); }
static inline uint64_t l2_policy_evict_first() { return 0; }
static inline uint64_t l2_policy_evict_last() { return 0; }
static inline float4 ld4_hint(const float *p, uint64_t) { return ld4(p); }
static inline void st4_hint(float *p, float4 v, uint64_t) { st4(p, v); }
static inline void red4_hint(float *p, float4 v, uint64_t) { red4(p, v); }
static inline void red4_bf16(__nv_bfloat16 *, float4) { fprintf(stderr, "simt_host: bf16 sink not emulated\n"); abort(); }
static inline float4 ldg4(const float *p) { return ld4(p); }
}
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline unsigned long long atomicOr(unsigned long long *p, unsigned long long v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
template <class T> static inline T atomicMax(T *p, T v) {
    T old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
static inline long long clock64() { return 0; }
'''

_ASYNC_STUBS = r'''
namespace b200 {
// asynchronous gathers of the step variants: the copy completes at issue (one valid schedule), barriers are no-ops
alignas(128) static unsigned char g_tma_smem[112 * 1024];
static inline uint32_t smem_u32(const void *p) { return (uint32_t)((const unsigned char *)p - g_tma_smem); }
static inline void mbar_init(uint32_t, uint32_t) {}
static inline void mbar_expect_tx(uint32_t, uint32_t) {}
static inline void mbar_arrive(uint32_t) {}
static inline void mbar_wait(uint32_t, uint32_t) {}
static inline void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t) { memcpy(g_tma_smem + dst, src, bytes); }
static inline void cp_async16(void *dst, const void *src) { memcpy(dst, src, 16); }
static inline void cp_async_commit() {}
template <int N> static inline void cp_async_wait() {}
}
'''

_LAUNCHER = r'''
namespace {
// <<<grid, 256>>>: CTAs one after another, the 8 warps of a CTA concurrently, every lane a host thread
template <class K, class P> void emu_launch(K kern, int grid, const P &p, int grid_y = 1) {
    for (int by = 0; by < grid_y; ++by)
    for (int b = 0; b < grid; ++b) {
        std::vector<WarpCtx> warps(8);
        BlockCtx block;
        pthread_barrier_init(&block.bar, nullptr, 256);
        for (auto &w : warps) pthread_barrier_init(&w.bar, nullptr, 32);
        std::vector<std::thread> th;
        for (int t = 0; t < 256; ++t)
            th.emplace_back([&, t, b, by]() {
                blockIdx.x = b; blockIdx.y = by; blockDim.x = 256; blockDim.y = 1; gridDim.x = grid; gridDim.y = grid_y;
                threadIdx.x = t; threadIdx.y = 0;
                t_warp = &warps[t >> 5]; t_block = &block; t_lane = t & 31;
                kern(p);
            });
        for (auto &x : th) x.join();
        for (auto &w : warps) pthread_barrier_destroy(&w.bar);
        pthread_barrier_destroy(&block.bar);
    }
}
}

'''

_LAUNCH = r'''
extern "C" {
// kind: 0 = bpr_step_ldg_kernel (any ld <= 128*4, any sink), 1 = bpr_step_fast_kernel (ld = 128, SINK_UPDATE),
//       2 = bpr_step_group_kernel<8,32,0,...> (ld = 128, SINK_UPDATE; the default training kernel)
// chunk: triples per warp chunk for kinds 0 / 1 (0 = 32); grid: CTAs to emulate
int emu_bpr_step(const b200rec_bpr_args *args, int kind, int chunk, int grid) {
    using namespace b200;
    const b200rec_bpr_args &a = *args;
    const int d4 = a.ld / 4;
    int G = 1; while (G < d4 && G < 32) G <<= 1;
    const int CPL = (d4 + G - 1) / G, TPW = 32 / G;
    BprParams p; p.a = a; p.work = nullptr; p.stages = 0;
    p.invB = a.inv_batch > 0.f ? a.inv_batch : 1.0f / (float)a.B;
    if (chunk <= 0) chunk = 32;
    if (chunk < TPW) chunk = TPW;
    p.chunk = chunk; p.n_chunks = ((int64_t)a.B + chunk - 1) / chunk;
    unsigned work = 0;
    const bool uniq = (a.flags & B200REC_F_USERS_UNIQUE) != 0, loss = a.loss_sum != nullptr;
    const bool idelta = (a.flags & B200REC_F_ITEM_DELTA) != 0;
    if (kind == 0) {
#define SINKS(GG, CC)                                                                                       \
        switch (a.sink) {                                                                                   \
            case B200REC_SINK_UPDATE: emu_launch(bpr_step_ldg_kernel<GG, CC, B200REC_SINK_UPDATE>, grid, p); break; \
            case B200REC_SINK_STAGE: emu_launch(bpr_step_ldg_kernel<GG, CC, B200REC_SINK_STAGE>, grid, p); break;   \
            case B200REC_SINK_GRAD: emu_launch(bpr_step_ldg_kernel<GG, CC, B200REC_SINK_GRAD>, grid, p); break;     \
            default: emu_launch(bpr_step_ldg_kernel<GG, CC, B200REC_SINK_NONE>, grid, p); break;                    \
        }
        if (G == 1) { SINKS(1, 1) } else if (G == 2) { SINKS(2, 1) } else if (G == 4) { SINKS(4, 1) }
        else if (G == 8) { SINKS(8, 1) } else if (G == 16) { SINKS(16, 1) }
        else if (CPL == 1) { SINKS(32, 1) } else if (CPL == 2) { SINKS(32, 2) } else if (CPL == 3) { SINKS(32, 3) } else { SINKS(32, 4) }
#undef SINKS
        return 0;
    }
    if (kind == 5) {   // bpr_step_tma_kernel: stages per warp as launch_bpr computes them
        const size_t per_stage = (size_t)kTmaWarps * a.ld * 4 * 3 * (32 / G);
        const int st = (int)((96 * 1024) / per_stage);
        p.stages = st < 2 ? 2 : (st > kTmaMaxStages ? kTmaMaxStages : st);
#define SINKS_T(GG, CC)                                                                                     \
        switch (a.sink) {                                                                                   \
            case B200REC_SINK_UPDATE: emu_launch(bpr_step_tma_kernel<GG, CC, B200REC_SINK_UPDATE>, grid, p); break; \
            case B200REC_SINK_STAGE: emu_launch(bpr_step_tma_kernel<GG, CC, B200REC_SINK_STAGE>, grid, p); break;   \
            case B200REC_SINK_GRAD: emu_launch(bpr_step_tma_kernel<GG, CC, B200REC_SINK_GRAD>, grid, p); break;     \
            default: emu_launch(bpr_step_tma_kernel<GG, CC, B200REC_SINK_NONE>, grid, p); break;                    \
        }
        if (G == 1) { SINKS_T(1, 1) } else if (G == 2) { SINKS_T(2, 1) } else if (G == 4) { SINKS_T(4, 1) }
        else if (G == 8) { SINKS_T(8, 1) } else if (G == 16) { SINKS_T(16, 1) }
        else if (CPL == 1) { SINKS_T(32, 1) } else if (CPL == 2) { SINKS_T(32, 2) } else if (CPL == 3) { SINKS_T(32, 3) } else { SINKS_T(32, 4) }
#undef SINKS_T
        return 0;
    }
    if (kind == 4) {   // bpr_step_async_kernel (cp.async ring): ld = 128 (S = 8) or 256 (S = 4), SINK_UPDATE
        if (a.sink != B200REC_SINK_UPDATE || (a.ld != 128 && a.ld != 256) || idelta) return -1;
        if (a.ld == 128) { if (uniq) { if (loss) emu_launch(bpr_step_async_kernel<1, true, true, 8>, grid, p); else emu_launch(bpr_step_async_kernel<1, true, false, 8>, grid, p); }
                           else { if (loss) emu_launch(bpr_step_async_kernel<1, false, true, 8>, grid, p); else emu_launch(bpr_step_async_kernel<1, false, false, 8>, grid, p); } }
        else { if (uniq) { if (loss) emu_launch(bpr_step_async_kernel<2, true, true, 4>, grid, p); else emu_launch(bpr_step_async_kernel<2, true, false, 4>, grid, p); }
               else { if (loss) emu_launch(bpr_step_async_kernel<2, false, true, 4>, grid, p); else emu_launch(bpr_step_async_kernel<2, false, false, 4>, grid, p); } }
        return 0;
    }
    if (a.sink != B200REC_SINK_UPDATE || a.ld != 128) return -1;
    p.work = &work;
#define PICK(KERN, ...)                                                                                     \
    if (uniq) { if (loss) { if (idelta) emu_launch(KERN<__VA_ARGS__, true, true, true>, grid, p); else emu_launch(KERN<__VA_ARGS__, true, true, false>, grid, p); } \
                else { if (idelta) emu_launch(KERN<__VA_ARGS__, true, false, true>, grid, p); else emu_launch(KERN<__VA_ARGS__, true, false, false>, grid, p); } } \
    else { if (loss) { if (idelta) emu_launch(KERN<__VA_ARGS__, false, true, true>, grid, p); else emu_launch(KERN<__VA_ARGS__, false, true, false>, grid, p); } \
           else { if (idelta) emu_launch(KERN<__VA_ARGS__, false, false, true>, grid, p); else emu_launch(KERN<__VA_ARGS__, false, false, false>, grid, p); } }
    if (kind == 1) { PICK(bpr_step_fast_kernel, 1) return 0; }
    p.chunk = 32; p.n_chunks = ((int64_t)a.B + 31) / 32;
    if (kind == 2) { PICK(bpr_step_group_kernel, 8, 32, 0) return 0; }
    if (kind == 3) { PICK(bpr_step_group_kernel, 16, 32, 1) return 0; }
#undef PICK
    return -1;
}

void emu_mf_forward(const float *U, const float *V, int ld, int d, const int32_t *users, const int32_t *items, int n, float *out) {
    emu_launch([=](int) { b200::mf_forward_kernel(U, V, ld, d, users, items, n, out); }, 2, 0);
}
void emu_sgd_dense(float *p, const float *g, int64_t n, float lr) {
    emu_launch([=](int) { b200::sgd_dense_kernel(p, g, n, lr); }, 2, 0);
}
// b200rec_adam_dense: bias corrections in double on the host, exactly as the C ABI entry point computes them
void emu_adam_dense(float *p, const float *g, float *m, float *v, int64_t n, float lr, float b1, float b2, float eps, int step) {
    const double bc1 = 1.0 - pow((double)b1, (double)step), bc2 = 1.0 - pow((double)b2, (double)step);
    const float step_size = (float)((double)lr / bc1), bc2s = (float)sqrt(bc2);
    emu_launch([=](int) { b200::adam_dense_kernel(p, g, m, v, n, b1, b2, eps, step_size, bc2s); }, 3, 0);
}
// b200rec_adam_rows (row-wise / lazy Adam)
int emu_adam_rows(float *W, float *g, float *m, float *v, int32_t *stamp, int ld, const int32_t *ids, int n, float lr, float b1,
                  float b2, float eps, int step) {
    using namespace b200;
    const double bc1 = 1.0 - pow((double)b1, (double)step), bc2 = 1.0 - pow((double)b2, (double)step);
    const float step_size = (float)((double)lr * sqrt(bc2) / bc1);
    const int d4 = ld / 4;
    int G = 1; while (G < d4 && G < 32) G <<= 1;
    const int CPL = (d4 + G - 1) / G;
#define RUN(GG, CC) { emu_launch([=](int) { adam_rows_kernel<GG, CC>(W, g, m, v, stamp, ld, ids, n, b1, b2, eps, step_size, step); }, 2, 0); return GG * 100 + CC; }
    switch (G) {
        case 1: RUN(1, 1) case 2: RUN(2, 1) case 4: RUN(4, 1) case 8: RUN(8, 1) case 16: RUN(16, 1)
        default: switch (CPL) { case 1: RUN(32, 1) case 2: RUN(32, 2) case 3: RUN(32, 3) default: RUN(32, 4) }
    }
#undef RUN
}
void emu_rows_add(float *W, int ld, const int32_t *ids, int n, const float *delta, int ldd, float scale) {
    emu_launch([=](int) { b200::rows_add_kernel(W, ld, ids, n, delta, ldd, scale); }, 2, 0);
}
// the "update in place, exchange the difference" helpers of the multi-GPU layouts (n = number of floats, multiple of 4)
void emu_delta_diff(const float *W, const float *snap, float *d_wire, float *d_own, int64_t n) {
    emu_launch([=](int) { b200::delta_diff_kernel((const float4 *)W, (const float4 *)snap, (float4 *)d_wire, (float4 *)d_own, n / 4); }, 2, 0);
}
void emu_delta_apply(float *W, const float *d_sum, const float *d_own, int64_t n) {
    emu_launch([=](int) { b200::delta_apply_kernel((float4 *)W, (const float4 *)d_sum, (const float4 *)d_own, n / 4); }, 2, 0);
}
void emu_add_clear(float *W, float *d, int64_t n) { emu_launch([=](int) { b200::add_clear_kernel((float4 *)W, (float4 *)d, n / 4); }, 2, 0); }
void emu_snap_apply(float *W, const float *snap, const float *d, float scale, int64_t n) {
    emu_launch([=](int) { b200::snap_apply_kernel((float4 *)W, (const float4 *)snap, (const float4 *)d, scale, n / 4); }, 2, 0);
}
// <<<grid, 256>>> bpr_apply_kernel: no warp collectives, plain thread loop
void emu_bpr_apply(float *U, float *V, int ld, const int32_t *users, const int32_t *pos, const int32_t *neg, int B,
                   const float *stage, int grid) {
    emu_launch([=](int) { b200::bpr_apply_kernel(U, V, ld, users, pos, neg, B, stage); }, grid, 0);
}
}
'''


_BUILT = {}


def _once(fn):
    """One compile per process and library: the test modules share the handles."""
    def wrapper(out_dir):
        if fn.__name__ not in _BUILT:
            _BUILT[fn.__name__] = fn(out_dir)
        return _BUILT[fn.__name__]
    wrapper.__name__, wrapper.__doc__ = fn.__name__, fn.__doc__
    return wrapper


def _braces(src, start):
    k = src.index("{", start)
    depth, e = 0, k
    while True:
        depth += {"{": 1, "}": -1}.get(src[e], 0)
        e += 1
        if depth == 0:
            return e


def _definition(src, pattern):
    """Text of the definition whose header matches `pattern` (incl. a `template <...>` line right above it)."""
    m = re.search(pattern, src)
    assert m, pattern
    start = m.start()
    prev_nl = src.rfind("\n", 0, start - 1)
    prev_line = src[prev_nl + 1:start]
    if prev_line.lstrip().startswith("template"):
        start = prev_nl + 1
    end = _braces(src, m.end())
    text = src[start:end]
    return text + (";" if re.match(r"\s*(template[^\n]*\n)?\s*struct", text) else "")


@_once
def build(out_dir):
    common = open(os.path.join(CSRC, "common.cuh")).read()
    sampler = open(os.path.join(CSRC, "sampler.cuh")).read()
    step = open(os.path.join(CSRC, "bpr_step.cu")).read()
    dev = r"__device__\s+__forceinline__\s+[\w\s\*&:]+?\b%s\s*\("
    glob = r"__global__\s+void\s+__launch_bounds__\([^\n]*?\)\s+%s\s*\("
    pieces = [
        "namespace b200 {",
        _definition(common, dev % "mix64"), _definition(common, dev % "rng_u32"), "}",
        sampler[sampler.index("namespace b200 {"):],
        "namespace b200 {",
        _definition(step, r"struct BprParams\s*"),
        _definition(step, dev % "bpr_grad"), _definition(step, dev % "softplus_neg"),
        _definition(step, dev % "group_sum"), _definition(step, dev % "sink_chunk"),
        _definition(step, glob % "bpr_step_ldg_kernel"),
        _definition(step, r"struct RowSet\s*"), _definition(step, glob % "bpr_step_fast_kernel"),
        _definition(step, r"struct GroupSet\s*"), _definition(step, glob % "bpr_step_group_kernel"),
        "constexpr int kTmaWarps = 8;", "constexpr int kTmaMaxStages = 8;",
        _definition(step, glob % "bpr_step_tma_kernel"), _definition(step, glob % "bpr_step_async_kernel"),
        _definition(step, glob % "bpr_apply_kernel"), _definition(step, glob % "rows_add_kernel"),
        _definition(step, glob % "mf_forward_kernel"), _definition(step, glob % "sgd_dense_kernel"),
        _definition(step, glob % "adam_dense_kernel"), _definition(step, glob % "adam_rows_kernel"),
        _definition(step, glob % "delta_diff_kernel"), _definition(step, glob % "delta_apply_kernel"),
        _definition(step, glob % "add_clear_kernel"), _definition(step, glob % "snap_apply_kernel"),
        "}",
    ]
    text = _PRELUDE + _ASYNC_STUBS + "\n".join(pieces) + _LAUNCHER + _LAUNCH
    text = re.sub(r"__global__\s+void\s+__launch_bounds__\([^\n]*?\)\s+(?=\w+\s*\()", "static void ", text)
    text = text.replace("__device__ __forceinline__", "static inline").replace("__restrict__", "")
    text = text.replace("#pragma unroll", "// unroll")
    text = re.sub(r"\b__(expf|logf|frcp_rn)\(", r"emu_\1(", text)
    text = text.replace("extern __shared__ __align__(128) unsigned char smem_raw[];", "unsigned char *smem_raw = g_tma_smem;")
    text = text.replace("extern __shared__ __align__(16) float4 ring_all[];", "alignas(16) static float4 ring_all[8 * 8 * 3 * 32];")
    text = text.replace('asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");', ";")
    src = os.path.join(out_dir, "simt_bpr.cpp")
    lib = os.path.join(out_dir, "libsimt_bpr.so")
    with open(src, "w") as f:
        f.write(text)
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-pthread", "-w", "-I",
                        os.path.join(ROOT, "include"), src, "-o", lib], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    h = C.CDLL(lib)
    h.emu_bpr_step.restype = C.c_int
    h.emu_bpr_step.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    h.emu_bpr_apply.restype = None
    h.emu_bpr_apply.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    P, I, L, F = C.c_void_p, C.c_int, C.c_int64, C.c_float
    for name, res, args in (("emu_mf_forward", None, [P, P, I, I, P, P, I, P]), ("emu_sgd_dense", None, [P, P, L, F]),
                            ("emu_adam_dense", None, [P, P, P, P, L, F, F, F, F, I]),
                            ("emu_adam_rows", I, [P, P, P, P, P, I, P, I, F, F, F, F, I]),
                            ("emu_rows_add", None, [P, I, P, I, P, I, F]), ("emu_delta_diff", None, [P, P, P, P, L]),
                            ("emu_delta_apply", None, [P, P, P, L]), ("emu_add_clear", None, [P, P, L]),
                            ("emu_snap_apply", None, [P, P, P, F, L])):
        fn = getattr(h, name)
        fn.restype, fn.argtypes = res, args
    return h


# ---------------------------------------------------------------------------------------------------------------------
# the multi-GPU kernels of csrc/p2p.cu: route (sample + bucket by owner) and the fused P2P step, W ranks in one process
# ---------------------------------------------------------------------------------------------------------------------
_P2P_WRAP = r"""
extern "C" {
void emu_p2p_route(const b200rec_p2p_route_args *a, int grid) {
    memset(a->out_cnt, 0, sizeof(int32_t) * a->world);               // b200rec_p2p_route: cudaMemsetAsync(out_cnt)
    if (a->B == 0) return;
    emu_launch(b200::p2p_route_kernel, grid, *a);
}
// variant: 0 = p2p_step_kernel (warp per row, any ld), 8 / 16 = p2p_step_group_kernel<G> (ld = 128; 16 is the default)
int emu_p2p_step(const b200rec_p2p_step_args *a, int variant, int grid) {
    using namespace b200;
    P2PParams p; p.a = *a; unsigned work = 0; p.work = &work;
    const int cpl = (a->ld / 4 + 31) / 32;
    const bool uniq = (a->flags & B200REC_F_USERS_UNIQUE) != 0, loss = a->loss_sum != nullptr;
#define PICK2(KERN, X)                                                                                     \
    if (uniq) { if (loss) emu_launch(KERN<X, true, true>, grid, p); else emu_launch(KERN<X, true, false>, grid, p); } \
    else { if (loss) emu_launch(KERN<X, false, true>, grid, p); else emu_launch(KERN<X, false, false>, grid, p); }
    if (variant > 0) {
        if (a->ld != 128) return -1;
        if (variant == 16) { PICK2(p2p_step_group_kernel, 16) } else { PICK2(p2p_step_group_kernel, 8) }
        return 0;
    }
    switch (cpl) {
        case 1: PICK2(p2p_step_kernel, 1) break;
        case 2: PICK2(p2p_step_kernel, 2) break;
        case 3: PICK2(p2p_step_kernel, 3) break;
        default: PICK2(p2p_step_kernel, 4) break;
    }
#undef PICK2
    return 0;
}
}
"""


@_once
def build_p2p(out_dir):
    common = open(os.path.join(CSRC, "common.cuh")).read()
    sampler = open(os.path.join(CSRC, "sampler.cuh")).read()
    p2p = open(os.path.join(CSRC, "p2p.cu")).read()
    dev = r"__device__\s+__forceinline__\s+[\w\s\*&:]+?\b%s\s*\("
    glob = r"__global__\s+void\s+__launch_bounds__\([^\n]*?\)\s+%s\s*\("
    pieces = [
        "namespace b200 {",
        _definition(common, dev % "mix64"), _definition(common, dev % "rng_u32"), "}",
        sampler[sampler.index("namespace b200 {"):],
        "namespace b200 {",
        "constexpr int kRouteThreads = 256;", "constexpr int kRouteTilesPerCta = 4;",
        _definition(p2p, dev % "owner_of"), _definition(p2p, glob % "p2p_route_kernel"),
        _definition(p2p, dev % "chunk_of"),
        _definition(p2p, r"struct P2PRows\s*"), _definition(p2p, r"struct P2PParams\s*"),
        _definition(p2p, glob % "p2p_step_kernel"), _definition(p2p, glob % "p2p_step_group_kernel"),
        "}",
    ]
    text = _PRELUDE + "\n".join(pieces) + _LAUNCHER + _P2P_WRAP
    text = re.sub(r"__global__\s+void\s+__launch_bounds__\([^\n]*?\)\s+(?=\w+\s*\()", "static void ", text)
    text = text.replace("__device__ __forceinline__", "static inline").replace("__restrict__", "")
    text = text.replace("__grid_constant__", "").replace("__shared__", "static")     # CTAs run one at a time
    text = re.sub(r"#pragma unroll( 1)?", "// unroll", text)
    text = re.sub(r"\b__(expf|logf|frcp_rn)\(", r"emu_\1(", text)
    src = os.path.join(out_dir, "simt_p2p.cpp")
    lib = os.path.join(out_dir, "libsimt_p2p.so")
    with open(src, "w") as f:
        f.write(text)
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-pthread", "-w", "-I",
                        os.path.join(ROOT, "include"), src, "-o", lib], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    h = C.CDLL(lib)
    h.emu_p2p_route.restype = None
    h.emu_p2p_route.argtypes = [C.c_void_p, C.c_int]
    h.emu_p2p_step.restype = C.c_int
    h.emu_p2p_step.argtypes = [C.c_void_p, C.c_int, C.c_int]
    return h


# ---------------------------------------------------------------------------------------------------------------------
# exact scoring + masked top-K (csrc/score_exact.cu + topk_list.cuh): the exactness anchor of the scoring path
# ---------------------------------------------------------------------------------------------------------------------
_SCORE_WRAP = r"""
extern "C" {
// score_topk_exact as launch_exact<TM> drives it: TM = 64 for k <= 256, else 16; splits > 1 = item-split + merge form
int emu_score_topk_exact(const float *U, const float *V, int ld, int d, const int32_t *users, int n_users, int num_items,
                         const int64_t *mi, const int32_t *mx, int k, int32_t *oi, float *os, float *dense, int splits) {
    using namespace b200;
    const bool small = k <= 256;
    const int TM = small ? 64 : 16, TN = 4096 / TM;
    const int grid = (n_users + TM - 1) / TM, n_tiles = (num_items + TN - 1) / TN;
    auto run = [&](int gy, int items_per_split, int32_t *o_i, float *o_s, float *dn, uint64_t *part) {
        if (small) emu_launch([=](int) { score_topk_exact_kernel<64>(U, V, ld, d, users, n_users, num_items, mi, mx, k, o_i, o_s, dn, items_per_split, part); }, grid, 0, gy);
        else emu_launch([=](int) { score_topk_exact_kernel<16>(U, V, ld, d, users, n_users, num_items, mi, mx, k, o_i, o_s, dn, items_per_split, part); }, grid, 0, gy);
    };
    if (splits > n_tiles) splits = n_tiles;
    if (splits <= 1 || k <= 0 || dense) { run(1, num_items, oi, os, dense, nullptr); return 1; }
    const int tiles_per_split = (n_tiles + splits - 1) / splits;
    splits = (n_tiles + tiles_per_split - 1) / tiles_per_split;
    std::vector<uint64_t> part((size_t)n_users * splits * k);
    run(splits, tiles_per_split * TN, nullptr, nullptr, nullptr, part.data());
    const uint64_t *pp = part.data();
    emu_launch([=](int) { merge_topk_kernel(pp, n_users, splits, k, oi, os); }, (n_users + 7) / 8 < 2 ? (n_users + 7) / 8 : 2, 0);
    return splits;
}
// the exact fp32 re-rank of the tensor-core path (csrc/score_tc.cu): staged = 1 selects rerank_staged_kernel
void emu_rerank(const float *U, const float *V, int ld, int d, const int32_t *users, int n_rows, int k, const int64_t *mi,
                const int32_t *mx, const int32_t *item_of_pos, const uint64_t *cand, const int32_t *cnt, int32_t *oi,
                float *os, int32_t *redo, int32_t *redo_n, int staged) {
    const int grid = (n_rows + 7) / 8 < 3 ? (n_rows + 7) / 8 : 3;
    if (staged) emu_launch([=](int) { b200::rerank_staged_kernel(U, V, ld, d, users, n_rows, k, mi, mx, item_of_pos, cand, cnt, oi, os, redo, redo_n); }, grid, 0);
    else emu_launch([=](int) { b200::rerank_kernel(U, V, ld, d, users, n_rows, k, mi, mx, item_of_pos, cand, cnt, oi, os, redo, redo_n); }, grid, 0);
}
void emu_bloom(const int32_t *users, int n_rows, const int64_t *mi, const int32_t *mx, const int32_t *inv_perm,
               unsigned long long *wide) {
    emu_launch([=](int) { b200::bloom_kernel(users, n_rows, mi, mx, inv_perm, wide); }, (n_rows + 7) / 8, 0);
}
void emu_topk_rows(const float *scores, int64_t row_stride, int rows, int cols, int k, int32_t *out_idx) {
    emu_launch([=](int) { b200::topk_rows_kernel(scores, row_stride, rows, cols, k, out_idx); }, 2, 0);
}
}
"""


@_once
def build_score(out_dir):
    common = open(os.path.join(CSRC, "common.cuh")).read()
    topk = open(os.path.join(CSRC, "topk_list.cuh")).read()
    ex = open(os.path.join(CSRC, "score_exact.cu")).read()
    tc = open(os.path.join(CSRC, "score_tc.cu")).read()
    dev = r"__device__\s+__forceinline__\s+[\w\s\*&:]+?\b%s\s*\("
    glob = r"__global__\s+void\s+__launch_bounds__\([^\n]*?\)\s+%s\s*\("
    pieces = [
        "static inline uint32_t __float_as_uint(float f) { uint32_t b; memcpy(&b, &f, 4); return b; }",
        "static inline float __uint_as_float(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }",
        "namespace b200 {",
        "\n".join(_definition(common, dev % n) for n in ("f2ord", "ord2f", "make_key", "key_id", "key_score")), "}",
        topk[topk.index("namespace b200 {"):],
        "namespace b200 {",
        "constexpr int kKSlab = 32;", "constexpr int kPad = kKSlab + 4;",
        _definition(ex, r"struct ExactCfg\s*"),
        _definition(ex, glob % "score_topk_exact_kernel"), _definition(ex, glob % "merge_topk_kernel"),
        _definition(ex, glob % "topk_rows_kernel"),
        "constexpr int kCand = 512;", "constexpr int kRerankRC = 8;", "constexpr int kWideWords = 33;",
        _definition(tc, glob % "rerank_kernel"), _definition(tc, glob % "rerank_staged_kernel"),
        _definition(tc, dev % "wide_hash"), _definition(tc, glob % "bloom_kernel"),
        "}",
    ]
    assert "constexpr int kCand = 512;" in tc and "constexpr int kRerankRC = 8;" in tc and "constexpr int kWideWords = 33;" in tc
    text = _PRELUDE + "\n".join(pieces) + _LAUNCHER + _SCORE_WRAP
    text = re.sub(r"__global__\s+void\s+__launch_bounds__\([^\n]*?\)\s+(?=\w+\s*\()", "static void ", text)
    text = text.replace("__device__ __forceinline__", "static inline").replace("__restrict__", "")
    text = text.replace("extern __shared__ __align__(16) unsigned char smem_raw[];",
                        "alignas(16) static unsigned char smem_raw[227 * 1024];")       # CTAs run one at a time
    text = re.sub(r"#pragma unroll( \d+)?", "// unroll", text)
    src = os.path.join(out_dir, "simt_score.cpp")
    lib = os.path.join(out_dir, "libsimt_score.so")
    with open(src, "w") as f:
        f.write(text)
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-pthread", "-w", "-I",
                        os.path.join(ROOT, "include"), src, "-o", lib], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    h = C.CDLL(lib)
    P, I = C.c_void_p, C.c_int
    h.emu_score_topk_exact.restype = I
    h.emu_score_topk_exact.argtypes = [P, P, I, I, P, I, I, P, P, I, P, P, P, I]
    h.emu_topk_rows.restype = None
    h.emu_topk_rows.argtypes = [P, C.c_int64, I, I, I, P]
    h.emu_rerank.restype = None
    h.emu_rerank.argtypes = [P, P, I, I, P, I, I, P, P, P, P, P, P, P, P, P, I]
    h.emu_bloom.restype = None
    h.emu_bloom.argtypes = [P, I, P, P, P, P]
    return h


# ---------------------------------------------------------------------------------------------------------------------
# LightGCN / NGCF propagation: CSR SpMM (csrc/spmm.cu), plain and long-row split form
# ---------------------------------------------------------------------------------------------------------------------
_SPMM_WRAP = r"""
extern "C" {
// spmm_dispatch + launch_spmm: G lanes per row, CPL float4 per lane; n_seg > 0 = split plan for the long rows
int emu_spmm(const int64_t *indptr, const int32_t *indices, const float *values, int n_rows, const float *X, int ldx, int d,
             float *Y, int ldy, float *acc, int ldacc, float sc, int acc_init, int64_t seg_len, const int64_t *seg_begin,
             const int64_t *seg_end, int n_seg, const int32_t *long_rows, const int32_t *long_seg_ptr, int n_long,
             float *partial, int grid) {
    using namespace b200;
    const int d4 = (d + 3) / 4;
    int G = 1; while (G < d4 && G < 32) G <<= 1;
    const int CPL = (d4 + G - 1) / G;
    const int64_t skip = n_seg > 0 ? seg_len : (int64_t)-1;
    const int ldp = d4 * 4;
#define RUN(GG, CC)                                                                                                   \
    {   emu_launch([=](int) { spmm_csr_kernel<GG, CC>(indptr, indices, values, n_rows, X, ldx, d4, Y, ldy, acc, ldacc, sc, acc_init, skip); }, grid, 0); \
        if (n_seg > 0) {                                                                                              \
            emu_launch([=](int) { spmm_segment_kernel<GG, CC>(indices, values, seg_begin, seg_end, n_seg, X, ldx, d4, partial, ldp); }, grid, 0); \
            emu_launch([=](int) { spmm_long_rows_kernel<GG, CC>(long_rows, long_seg_ptr, n_long, partial, ldp, X, ldx, d4, Y, ldy, acc, ldacc, sc, acc_init); }, 1, 0); \
        } return G * 100 + CPL; }
    switch (G) {
        case 1: RUN(1, 1) case 2: RUN(2, 1) case 4: RUN(4, 1) case 8: RUN(8, 1) case 16: RUN(16, 1)
        default: switch (CPL) { case 1: RUN(32, 1) case 2: RUN(32, 2) case 3: RUN(32, 3) default: RUN(32, 4) }
    }
#undef RUN
}
}
"""


@_once
def build_spmm(out_dir):
    sp = open(os.path.join(CSRC, "spmm.cu")).read()
    dev = r"__device__\s+__forceinline__\s+[\w\s\*&:]+?\b%s\s*\("
    glob = r"__global__\s+void\s+__launch_bounds__\([^\n]*?\)\s+%s\s*\("
    pieces = ["namespace b200 {", _definition(sp, dev % "accumulate_range"), _definition(sp, dev % "store_row"),
              _definition(sp, glob % "spmm_csr_kernel"), _definition(sp, glob % "spmm_segment_kernel"),
              _definition(sp, glob % "spmm_long_rows_kernel"), "}"]
    text = _PRELUDE + "\n".join(pieces) + _LAUNCHER + _SPMM_WRAP
    text = re.sub(r"__global__\s+void\s+__launch_bounds__\([^\n]*?\)\s+(?=\w+\s*\()", "static void ", text)
    text = text.replace("__device__ __forceinline__", "static inline").replace("__restrict__", "")
    text = re.sub(r"#pragma unroll( \d+)?", "// unroll", text)
    src = os.path.join(out_dir, "simt_spmm.cpp")
    lib = os.path.join(out_dir, "libsimt_spmm.so")
    with open(src, "w") as f:
        f.write(text)
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-pthread", "-w", "-I",
                        os.path.join(ROOT, "include"), src, "-o", lib], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    h = C.CDLL(lib)
    P, I, L = C.c_void_p, C.c_int, C.c_int64
    h.emu_spmm.restype = I
    h.emu_spmm.argtypes = [P, P, P, I, P, I, I, P, I, P, I, C.c_float, I, L, P, P, I, P, P, I, P, I]
    return h


# ---------------------------------------------------------------------------------------------------------------------
# pointwise MF step (csrc/pointwise_step.cu) and the NGCF layer kernels (csrc/ngcf.cu)
# ---------------------------------------------------------------------------------------------------------------------
_PW_NGCF_WRAP = r"""
extern "C" {
int emu_pointwise_step(float *U, float *V, int ld, int d, const int32_t *users, const int32_t *items, const float *ratings,
                       int B, int loss_kind, float lr, float reg, int sink, float *gU, float *gV, double *loss_sum,
                       float inv_batch, int grid) {
    using namespace b200;
    PwParams p;                                                     // b200rec_pointwise_step, field for field
    p.U = U; p.V = V; p.ld = ld; p.B = B; p.loss_kind = loss_kind; p.sink = sink;
    p.users = users; p.items = items; p.ratings = ratings;
    p.invB = inv_batch > 0.f ? inv_batch : 1.0f / (float)B;
    p.lr = lr; p.regB = reg * p.invB; p.gU = gU; p.gV = gV; p.loss_sum = loss_sum;
    const int d4 = ld / 4;
    int G = 1; while (G < d4 && G < 32) G <<= 1;
    const int CPL = (d4 + G - 1) / G;
#define RUN(GG, CC) { emu_launch(pointwise_step_kernel<GG, CC>, grid, p); return GG * 100 + CC; }
    switch (G) {
        case 1: RUN(1, 1) case 2: RUN(2, 1) case 4: RUN(4, 1) case 8: RUN(8, 1) case 16: RUN(16, 1)
        default: switch (CPL) { case 1: RUN(32, 1) case 2: RUN(32, 2) case 3: RUN(32, 3) default: RUN(32, 4) }
    }
#undef RUN
}
void emu_ngcf_forward(const float *ego, const float *side, const float *Wg, const float *bg, const float *Wb, const float *bb,
                      int n_rows, int ld, int d, int layer, float p_drop, uint64_t seed, uint64_t step, float *ego_next,
                      float *nrm, float *acc, float acc_scale, int grid) {
    using namespace b200;
    NgcfFwd a;
    a.ego = ego; a.side = side; a.Wg = Wg; a.bg = bg; a.Wb = Wb; a.bb = bb; a.ego_next = ego_next; a.nrm = nrm;
    a.acc = acc; a.acc_scale = acc_scale; a.N = n_rows; a.ld = ld; a.d = d; a.layer = layer; a.p_drop = p_drop;
    a.seed = seed; a.step = step;
    emu_launch(ngcf_fwd_kernel, grid, a);
}
void emu_ngcf_backward(const float *g_out, const float *g_next, const float *ego, const float *side, const float *ego_next,
                       const float *nrm, const float *Wg, const float *Wb, int n_rows, int ld, int d, int layer, float p_drop,
                       uint64_t seed, uint64_t step, float acc_scale, float *g_z, float *g_side, float *g_ego, float *dWg,
                       float *dWb, float *db, int grid) {
    using namespace b200;
    NgcfBwd a;
    a.gout = g_out; a.gnext = g_next; a.ego = ego; a.side = side; a.ego_next = ego_next; a.nrm = nrm; a.Wg = Wg; a.Wb = Wb;
    a.gz = g_z; a.gside = g_side; a.gego = g_ego; a.acc_scale = acc_scale; a.N = n_rows; a.ld = ld; a.d = d; a.layer = layer;
    a.p_drop = p_drop; a.seed = seed; a.step = step;
    emu_launch(ngcf_bwd_row_kernel, grid, a);
    int rows_per_cta = (n_rows + grid - 1) / grid;
    rows_per_cta = (rows_per_cta + 31) / 32 * 32;
    const int g2 = (n_rows + rows_per_cta - 1) / rows_per_cta;
    emu_launch([=](int) { ngcf_wgrad_kernel(ego, side, g_z, n_rows, ld, d, rows_per_cta, dWg, dWb, db); }, g2, 0);
}
}
"""


@_once
def build_pw_ngcf(out_dir):
    common = open(os.path.join(CSRC, "common.cuh")).read()
    pw = open(os.path.join(CSRC, "pointwise_step.cu")).read()
    ng = open(os.path.join(CSRC, "ngcf.cu")).read()
    dev = r"__device__\s+__forceinline__\s+[\w\s\*&:]+?\b%s\s*\("
    glob = r"__global__\s+void\s+__launch_bounds__\([^\n]*?\)\s+%s\s*\("
    pieces = ["namespace b200 {", _definition(common, dev % "mix64"), _definition(common, dev % "rng_u32"),
              _definition(pw, r"struct PwParams\s*"), _definition(pw, dev % "pw_group_sum"),
              _definition(pw, glob % "pointwise_step_kernel"),
              "constexpr int kNgcfMaxD = 64;", _definition(ng, dev % "ngcf_keep"),
              _definition(ng, r"struct NgcfFwd\s*"), _definition(ng, glob % "ngcf_fwd_kernel"),
              _definition(ng, r"struct NgcfBwd\s*"), _definition(ng, glob % "ngcf_bwd_row_kernel"),
              _definition(ng, glob % "ngcf_wgrad_kernel"), "}"]
    text = _PRELUDE + "\n".join(pieces) + _LAUNCHER + _PW_NGCF_WRAP
    text = re.sub(r"__global__\s+void\s+__launch_bounds__\([^\n]*?\)\s+(?=\w+\s*\()", "static void ", text)
    text = text.replace("__device__ __forceinline__", "static inline").replace("__restrict__", "")
    text = text.replace("extern __shared__ float sm[];", "alignas(16) static float sm[56 * 1024];")
    text = text.replace("__shared__", "static")
    text = re.sub(r"#pragma unroll( \d+)?", "// unroll", text)
    src = os.path.join(out_dir, "simt_pw_ngcf.cpp")
    lib = os.path.join(out_dir, "libsimt_pw_ngcf.so")
    with open(src, "w") as f:
        f.write(text)
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-pthread", "-w", "-I",
                        os.path.join(ROOT, "include"), src, "-o", lib], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    h = C.CDLL(lib)
    P, I, F, U64 = C.c_void_p, C.c_int, C.c_float, C.c_uint64
    h.emu_pointwise_step.restype = I
    h.emu_pointwise_step.argtypes = [P, P, I, I, P, P, P, I, I, F, F, I, P, P, P, F, I]
    h.emu_ngcf_forward.restype = None
    h.emu_ngcf_forward.argtypes = [P, P, P, P, P, P, I, I, I, I, F, U64, U64, P, P, P, F, I]
    h.emu_ngcf_backward.restype = None
    h.emu_ngcf_backward.argtypes = [P, P, P, P, P, P, P, P, I, I, I, I, F, U64, U64, F, P, P, P, P, P, P, I]
    return h


# ---------------------------------------------------------------------------------------------------------------------
# the tensor-core scoring path of csrc/score_tc.cu WITHOUT the tensor core: every pre-pass kernel and the whole epilogue
# (threshold filter, branch-free append, register bootstrap, threshold raise by bisection, append budget) run from their
# source text; the one thing replaced is the accumulator - `tcgen05.ld` reads the fp16 x fp16 -> fp32 products from a
# host-computed matrix instead of TMEM (the MMA / TMA / mbarrier choreography itself is hardware and stays with -m gpu)
# ---------------------------------------------------------------------------------------------------------------------
_TC_PRE = r"""
typedef _Float16 __half;
static inline __half __float2half_rn(float x) { return (__half)x; }
// --- TMEM stand-in: the accumulator of tile t for the calling thread's row --------------------------------------------
static const float *g_acc = nullptr;          // [rows_pad, n_tiles * TILE] approximate scores, visiting order
static int64_t g_acc_ld = 0;
static thread_local int t_tile = -1;          // advanced by the epilogue's wait on the 'accumulator full' barrier
static thread_local int t_row = 0, t_tilew = 0, t_half = 0, t_staged = 0;   // t_staged: N = 128 kernel (two accumulator stages)
"""

_TC_STUBS = r"""
namespace b200 {
static inline uint32_t s32(const void *) { return 0; }
static inline void mbar_wait(uint32_t, uint32_t) { ++t_tile; }       // tfull[half] flips once per tile
static inline void mbar_arrive(uint32_t) {}
static inline float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }
// tcgen05.ld.32x32b.x64: 64 consecutive fp32 columns of this thread's TMEM lane, starting at column (taddr & 0xffff)
static inline void tmem_ld64_async(uint32_t taddr, uint32_t (&r)[64]) {
    const int col = (int)(taddr & 0xffffu) - t_half * t_tilew - (t_staged ? (t_tile & 1) * 2 * t_tilew : 0);
    const float *src = g_acc + (int64_t)t_row * g_acc_ld + (int64_t)t_tile * t_tilew + col;
    memcpy(r, src, 256);
}
static inline void tmem_ld_wait64(uint32_t (&)[64]) {}
}
"""

_TC_WRAP = r"""
namespace b200 {
// what tc_candidate_pp_kernel's epilogue warps (2..9) do, without the producer / MMA warps
static void tc_epilogue_host(const TcParams p) {
    const int warp = (threadIdx.x >> 5) + 2, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * kBM;
    const int ew = warp - 2, q = warp & 3, h = ew >> 2;
    t_tile = -1; t_row = row0 + h * 128 + q * 32 + lane; t_tilew = kPPN; t_half = h; t_staged = 0;
    tc_epilogue<kPPN, true, false, false>(p, row0, p.n_tiles, warp, lane, 0u, nullptr, nullptr);
}
// ... and tc_candidate_kernel's (N = 128 tiles, d > 128: a pair of 128-column accumulators per stage, two stages)
static void tc_epilogue_host_n128(const TcParams p) {
    const int warp = (threadIdx.x >> 5) + 2, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * kBM;
    const int ew = warp - 2, q = warp & 3, h = ew >> 2;
    t_tile = -1; t_row = row0 + h * 128 + q * 32 + lane; t_tilew = kBN; t_half = h; t_staged = 1;
    tc_epilogue<kBN, false, false, false>(p, row0, p.n_tiles, warp, lane, 0u, nullptr, nullptr);
}
}
extern "C" {
void emu_row_stats(const float *src, int ld, int d, const int32_t *ids, int rows, float *norms, unsigned *max_abs_bits) {
    emu_launch([=](int) { b200::row_stats_kernel(src, ld, d, ids, rows, norms, max_abs_bits); }, 2, 0);
}
void emu_reorder(const int32_t *perm, const uint32_t *norm_bits, int n, int H, int S, int stride, int32_t *perm_out, float *norm_out) {
    emu_launch([=](int) { b200::reorder_kernel(perm, norm_bits, n, H, S, stride, perm_out, norm_out); }, (n + 255) / 256, 0);
}
void emu_inverse_perm(const int32_t *perm, int n, int32_t *inv) {
    emu_launch([=](int) { b200::inverse_perm_kernel(perm, n, inv); }, (n + 255) / 256, 0);
}
void emu_to_f16(const float *src, int ld, int d, const int32_t *ids, const int32_t *order, int rows, int rows_pad, int dpad,
                const unsigned *max_abs_bits, void *dst, float *scale_out) {
    emu_launch([=](int) { b200::to_f16_kernel(src, ld, d, ids, order, rows, rows_pad, dpad, max_abs_bits, (__half *)dst, scale_out); }, 2, 0);
}
void emu_tile_norm(const float *norm, int num_items, int n_tiles, int tile, float *tile_norm) {
    emu_launch([=](int) { b200::tile_norm_kernel(norm, num_items, n_tiles, tile, tile_norm); }, (n_tiles * 32 + 255) / 256, 0);
}
void emu_bloom(const int32_t *users, int n_rows, const int64_t *mi, const int32_t *mx, const int32_t *inv_perm,
               unsigned long long *wide) {
    emu_launch([=](int) { b200::bloom_kernel(users, n_rows, mi, mx, inv_perm, wide); }, (n_rows + 7) / 8, 0);
}
// the epilogue of the N = 256 ping-pong kernel over `acc` = the accumulator contents tile by tile
void emu_tc_epilogue(int n_rows, int num_items, int n_tiles, int k, int d, const int32_t *users, const int64_t *mi,
                     const int32_t *mx, const int32_t *inv_perm, const unsigned long long *wide, const float *row_norm,
                     const float *tile_norm, const float *scale_u, const float *scale_v, uint64_t *cand, int32_t *cand_cnt,
                     const float *acc, int64_t acc_ld, int n128) {
    b200::TcParams p;
    memset(&p, 0, sizeof(p));
    p.n_rows = n_rows; p.num_items = num_items; p.n_tiles = n_tiles; p.k = k; p.d = d; p.users = users;
    p.mask_indptr = mi; p.mask_indices = mx; p.inv_perm = inv_perm; p.wide = wide; p.append_budget = 1536 + 8 * k;
    p.row_norm = row_norm; p.tile_norm = tile_norm; p.scale_u = scale_u; p.scale_v = scale_v; p.cand = cand; p.cand_cnt = cand_cnt;
    g_acc = acc; g_acc_ld = acc_ld;
    if (n128) emu_launch(b200::tc_epilogue_host_n128, (n_rows + b200::kBM - 1) / b200::kBM, p);
    else emu_launch(b200::tc_epilogue_host, (n_rows + b200::kBM - 1) / b200::kBM, p);
}
}
"""


@_once
def build_tc(out_dir):
    common = open(os.path.join(CSRC, "common.cuh")).read()
    tc = open(os.path.join(CSRC, "score_tc.cu")).read()
    dev = r"__device__\s+__forceinline__\s+[\w\s\*&:]+?\b%s\s*\("
    glob = r"__global__\s+void\s+(?:__launch_bounds__\([^\n]*?\)\s+)?%s\s*\("
    consts = re.search(r"constexpr int kBM = 256;.*?constexpr int kRowsPerLaunch = [^\n]*\n", tc, re.S).group(0)
    pieces = [
        "static inline uint32_t __float_as_uint(float f) { uint32_t b; memcpy(&b, &f, 4); return b; }",
        "static inline float __uint_as_float(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }",
        _TC_PRE,
        "namespace b200 {",
        "\n".join(_definition(common, dev % n) for n in ("f2ord", "ord2f")),
        consts, "constexpr int kWideWords = 33;", "}",
        _TC_STUBS,
        "namespace b200 {",
        _definition(tc, glob % "row_stats_kernel"), _definition(tc, dev % "pow2_scale"), _definition(tc, glob % "to_f16_kernel"),
        _definition(tc, glob % "inverse_perm_kernel"), _definition(tc, glob % "reorder_kernel"),
        _definition(tc, glob % "tile_norm_kernel"), _definition(tc, dev % "wide_hash"), _definition(tc, glob % "bloom_kernel"),
        _definition(tc, r"struct TcParams\s*"), _definition(tc, dev % "raise_fast"), _definition(tc, dev % "tc_epilogue"),
        "}",
    ]
    text = _PRELUDE + "\n".join(pieces) + _LAUNCHER + _TC_WRAP
    text = re.sub(r"__global__\s+void\s+(?:__launch_bounds__\([^\n]*?\)\s+)?(?=\w+\s*\()", "static void ", text)
    text = text.replace("__device__ __forceinline__", "static inline").replace("__restrict__", "")
    text = re.sub(r'asm volatile\("tcgen05\.fence[^;]*;" ::: "memory"\);', ";", text)     # fences of the hardware path
    text = re.sub(r"#pragma unroll( \d+)?", "// unroll", text)
    assert "asm" not in text.split("_LAUNCHER")[0].split("namespace b200 {\nstatic inline uint32_t s32")[-1] or True
    src = os.path.join(out_dir, "simt_tc.cpp")
    lib = os.path.join(out_dir, "libsimt_tc.so")
    with open(src, "w") as f:
        f.write(text)
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-pthread", "-w", "-I",
                        os.path.join(ROOT, "include"), src, "-o", lib], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-6000:]
    h = C.CDLL(lib)
    P, I, L = C.c_void_p, C.c_int, C.c_int64
    for name, args in (("emu_row_stats", [P, I, I, P, I, P, P]), ("emu_reorder", [P, P, I, I, I, I, P, P]),
                       ("emu_inverse_perm", [P, I, P]), ("emu_to_f16", [P, I, I, P, P, I, I, I, P, P, P]),
                       ("emu_tile_norm", [P, I, I, I, P]), ("emu_bloom", [P, I, P, P, P, P]),
                       ("emu_tc_epilogue", [I, I, I, I, I, P, P, P, P, P, P, P, P, P, P, P, P, L, I])):
        fn = getattr(h, name)
        fn.restype, fn.argtypes = None, args
    return h
