// Shared helpers for libb200rec (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/b200rec.h"

namespace b200 {

// ---- error plumbing -------------------------------------------------------
void set_error(const char *fmt, ...);
void count_launch(int n = 1);

#define B200_REQUIRE(cond, code, ...)                 \
    do {                                              \
        if (!(cond)) {                                \
            ::b200::set_error(__VA_ARGS__);           \
            return (code);                            \
        }                                             \
    } while (0)

#define B200_CUDA(call)                                                                   \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            ::b200::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,                \
                              cudaGetErrorString(e__));                                   \
            return B200REC_ECUDA;                                                         \
        }                                                                                 \
    } while (0)

#define B200_LAUNCH_CHECK()                                                               \
    do {                                                                                  \
        cudaError_t e__ = cudaGetLastError();                                             \
        if (e__ != cudaSuccess) {                                                         \
            ::b200::set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__,            \
                              cudaGetErrorString(e__));                                   \
            return B200REC_ECUDA;                                                         \
        }                                                                                 \
        ::b200::count_launch();                                                           \
    } while (0)

int sm_count();  // SMs of the current device (cached)

// ---- device helpers -------------------------------------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// counter RNG keyed by (seed, step, triple index, draw); mirrored on the host by
// oracle/bpr_oracle.py::rng_u32 for the parity tests
__device__ __forceinline__ uint32_t rng_u32(uint64_t seed, uint64_t step, uint64_t idx, uint32_t draw) {
    uint64_t k = mix64(seed ^ mix64(step));
    k = mix64(k ^ ((idx << 8) | (uint64_t)(draw & 0xFFu)));
    return (uint32_t)(k >> 32);
}

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
// 16-byte vector reduction into global memory (REDG.E.ADD.F32x4 on sm_100a)
__device__ __forceinline__ void red4(float *p, float4 v) {
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}

// L2 eviction-priority policies (createpolicy) and 16-byte accesses that carry one: the fused BPR step streams
// 1 GB of user rows per launch through an L2 that should keep the 51 MB item table resident
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float4 ld4_hint(const float *p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st4_hint(float *p, float4 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void red4_hint(float *p, float4 v, uint64_t pol) {
    asm volatile("red.relaxed.gpu.global.add.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol)
                 : "memory");
}

// order-preserving float -> uint32 (larger float -> larger key); -inf is the smallest finite key
__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t b = __float_as_uint(f);
    return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t k) {
    uint32_t b = k ^ ((k >> 31) ? 0x80000000u : 0xFFFFFFFFu);
    return __uint_as_float(b);
}
// 64-bit ranking key: larger key == better == (score desc, id asc)
__device__ __forceinline__ uint64_t make_key(float s, int32_t id) {
    return ((uint64_t)f2ord(s) << 32) | (uint32_t)(~(uint32_t)id);
}
__device__ __forceinline__ int32_t key_id(uint64_t k) { return (int32_t)(~(uint32_t)(k & 0xFFFFFFFFu)); }
__device__ __forceinline__ float key_score(uint64_t k) { return ord2f((uint32_t)(k >> 32)); }

}  // namespace b200
