"""Test infrastructure: compile the INTEGER / INDEX device functions of the CUDA sources for the HOST (g++), straight
from their source text, so that `-m "not gpu"` tests can run the very code the kernels execute against the Python
mirrors in oracle/.  Nothing here is shipped or imported by the product.

What is taken (verbatim, with `__device__ __forceinline__` rewritten to `static inline`):
  csrc/common.cuh   mix64, rng_u32, f2ord, ord2f, make_key, key_id, key_score
  csrc/sampler.cuh  everything (sample_pos, sample_neg, sample_neg2, fetch_triple)
  csrc/metrics.cu   in_truth, holdout_kernel, loo_kernel (one thread per user: the host driver loops over the thread ids)
  csrc/p2p.cu       owner_of (shard of an item id), chunk_of (work order of the fused P2P step)
  csrc/score_tc.cu  pow2_scale (the exact power-of-two rescale in front of the fp16 candidate pass), reorder_kernel (the
                    visiting order head | stratified sample | rest)
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "recsys_pytorch_b200", "csrc")

_WRAP = r'''
extern "C" {
// fetch_triple for t = 0..B-1 exactly as a kernel lane calls it; valid_out[t] = 1 when the triple is processed
void host_fetch_triples(const b200rec_bpr_args *a, int *valid_out, int *u_out, int *i_out, int *j_out) {
    for (int64_t t = 0; t < a->B; ++t) {
        bool valid = true; int u, i, j;
        b200::fetch_triple(*a, t, valid, u, i, j);
        valid_out[t] = valid ? 1 : 0; u_out[t] = u; i_out[t] = i; j_out[t] = j;
    }
}
int host_sample_neg2(const int32_t *row, uint32_t deg, uint32_t cnt0, uint32_t lo1, uint32_t cnt1, uint64_t seed,
                     uint64_t step, uint64_t t, int *j) {
    return b200::sample_neg2(row, deg, cnt0, lo1, cnt1, seed, step, t, *j) ? 1 : 0;
}
uint32_t host_rng_u32(uint64_t seed, uint64_t step, uint64_t idx, uint32_t draw) { return b200::rng_u32(seed, step, idx, draw); }
uint64_t host_make_key(float s, int32_t id) { return b200::make_key(s, id); }
int32_t host_key_id(uint64_t k) { return b200::key_id(k); }
float host_key_score(uint64_t k) { return b200::key_score(k); }
}
'''


def _function(src, name):
    """Source text of `__device__ __forceinline__ <ret> name(...) { ... }` (brace matched)."""
    m = re.search(r"__device__\s+__forceinline__\s+[\w\s\*&:]+?\b%s\s*\(" % re.escape(name), src)
    assert m, name
    k = src.index("{", m.end())
    depth, e = 0, k
    while True:
        depth += {"{": 1, "}": -1}.get(src[e], 0)
        e += 1
        if depth == 0:
            break
    return src[m.start():e]


def build(out_dir):
    """Returns a ctypes handle of the host build."""
    common = open(os.path.join(CSRC, "common.cuh")).read()
    sampler = open(os.path.join(CSRC, "sampler.cuh")).read()
    body = sampler[sampler.index("namespace b200 {"):]                     # drop the pragma / include lines
    helpers = "\n".join(_function(common, n) for n in ("mix64", "rng_u32", "f2ord", "ord2f", "make_key", "key_id", "key_score"))
    text = "\n".join([
        "#include <stdint.h>", "#include <string.h>", '#include "b200rec.h"',
        "static inline uint32_t __float_as_uint(float f) { uint32_t b; memcpy(&b, &f, 4); return b; }",
        "static inline float __uint_as_float(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }",
        "namespace b200 {", helpers, "}", body, _WRAP]).replace("__device__ __forceinline__", "static inline")
    src = os.path.join(out_dir, "host_device_code.cpp")
    lib = os.path.join(out_dir, "libhost_device_code.so")
    with open(src, "w") as f:
        f.write(text)
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), src, "-o", lib],
                   check=True)
    h = C.CDLL(lib)
    h.host_rng_u32.restype = C.c_uint32
    h.host_rng_u32.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32]
    h.host_make_key.restype = C.c_uint64
    h.host_make_key.argtypes = [C.c_float, C.c_int32]
    h.host_key_id.restype = C.c_int32
    h.host_key_id.argtypes = [C.c_uint64]
    h.host_key_score.restype = C.c_float
    h.host_key_score.argtypes = [C.c_uint64]
    h.host_sample_neg2.restype = C.c_int
    h.host_sample_neg2.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64,
                                   C.c_uint64, C.POINTER(C.c_int)]
    h.host_fetch_triples.restype = None
    return h


_METRICS_WRAP = r"""
extern "C" {
void host_holdout(const int32_t *topk, int n, int max_k, const int32_t *row_ids, const int64_t *tptr, const int32_t *tidx,
                  const int *Ks, int K_len, float *out) {
    for (int i = 0; i < b200::kMaxK + 2; ++i) b200::c_inv_log2[i] = 1.0 / log2((double)(i + 2));   // upload_tables()
    b200::KsArg ks; ks.n = K_len; for (int j = 0; j < 64; ++j) ks.v[j] = j < K_len ? Ks[j] : 0;     // make_ks()
    blockDim.x = 128;
    for (int b = 0; b * 128 < n; ++b) for (int t = 0; t < 128; ++t) {
        blockIdx.x = b; threadIdx.x = t;
        b200::holdout_kernel(topk, n, max_k, row_ids, tptr, tidx, ks, out);
    }
}
void host_loo(const int32_t *topk, int n, int max_k, const int32_t *row_ids, const int64_t *tptr, const int32_t *tidx,
              const int *Ks, int K_len, float *out) {
    for (int i = 0; i < b200::kMaxK + 2; ++i) b200::c_inv_log2[i] = 1.0 / log2((double)(i + 2));
    b200::KsArg ks; ks.n = K_len; for (int j = 0; j < 64; ++j) ks.v[j] = j < K_len ? Ks[j] : 0;
    blockDim.x = 128;
    for (int b = 0; b * 128 < n; ++b) for (int t = 0; t < 128; ++t) {
        blockIdx.x = b; threadIdx.x = t;
        b200::loo_kernel(topk, n, max_k, row_ids, tptr, tidx, ks, out);
    }
}
}
"""


def _kernel(src, name):
    """Source text of `__global__ void __launch_bounds__(..) name(...) { ... }` as a plain function."""
    m = re.search(r"__global__\s+void\s+(?:__launch_bounds__\(\d+\)\s+)?%s\s*\(" % re.escape(name), src)
    assert m, name
    k = src.index("{", m.end())
    depth, e = 0, k
    while True:
        depth += {"{": 1, "}": -1}.get(src[e], 0)
        e += 1
        if depth == 0:
            break
    return re.sub(r"__global__\s+void\s+(?:__launch_bounds__\(\d+\))?", "static void ", src[m.start():e], count=1)


def build_metrics(out_dir):
    """Host build of the metric kernels of csrc/metrics.cu; returns a ctypes handle."""
    src_cu = open(os.path.join(CSRC, "metrics.cu")).read()
    ks_arg = re.search(r"struct KsArg \{.*?\};", src_cu, re.S).group(0)
    text = "\n".join([
        "#include <stdint.h>", "#include <math.h>",
        "struct Dim3 { int x; }; static thread_local Dim3 blockIdx, blockDim, threadIdx;",
        "namespace b200 {", "constexpr int kMaxK = 1024;", "static double c_inv_log2[kMaxK + 2];", ks_arg,
        _function(src_cu, "in_truth"), _kernel(src_cu, "holdout_kernel"), _kernel(src_cu, "loo_kernel"), "}",
        _METRICS_WRAP]).replace("__device__ __forceinline__", "static inline").replace("__restrict__", "")
    src = os.path.join(out_dir, "host_metrics.cpp")
    lib = os.path.join(out_dir, "libhost_metrics.so")
    with open(src, "w") as f:
        f.write(text)
    # -ffp-contract=off / no fast-math: the arithmetic must stay the IEEE sequence the source spells out
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", src, "-o", lib], check=True)
    return C.CDLL(lib)


_MISC_WRAP = r"""
extern "C" {
int host_owner_of(const int32_t *bounds, int world, int id) { return b200::owner_of(bounds, world, id); }
void host_chunk_of(int c, int W, int m, const int *pref, int round_robin, int *k, int *off) {
    b200::chunk_of(c, W, m, pref, round_robin != 0, *k, *off);
}
float host_pow2_scale(unsigned bits) { return b200::pow2_scale(bits); }
// reorder_kernel<<<(n + 255) / 256, 256>>> thread by thread
void host_reorder(const int32_t *perm, const uint32_t *norm_bits, int n, int H, int S, int stride, int32_t *perm_out,
                  float *norm_out) {
    blockDim.x = 256;
    for (int b = 0; b * 256 < n; ++b) for (int t = 0; t < 256; ++t) {
        blockIdx.x = b; threadIdx.x = t;
        b200::reorder_kernel(perm, norm_bits, n, H, S, stride, perm_out, norm_out);
    }
}
}
"""


def build_misc(out_dir):
    p2p = open(os.path.join(CSRC, "p2p.cu")).read()
    tc = open(os.path.join(CSRC, "score_tc.cu")).read()
    text = "\n".join([
        "#include <stdint.h>", "#include <string.h>", "#include <math.h>",
        "static inline float __uint_as_float(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }",
        "struct Dim3 { int x; }; static thread_local Dim3 blockIdx, blockDim, threadIdx;",
        "namespace b200 {", _function(p2p, "owner_of"), _function(p2p, "chunk_of"), _function(tc, "pow2_scale"),
        _kernel(tc, "reorder_kernel"), "}",
        _MISC_WRAP]).replace("__device__ __forceinline__", "static inline").replace("#pragma unroll 1", "").replace("__restrict__", "")
    src = os.path.join(out_dir, "host_misc.cpp")
    lib = os.path.join(out_dir, "libhost_misc.so")
    with open(src, "w") as f:
        f.write(text)
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", src, "-o", lib], check=True)
    h = C.CDLL(lib)
    h.host_pow2_scale.restype = C.c_float
    h.host_pow2_scale.argtypes = [C.c_uint32]
    return h
