// Tensor-core scoring + top-K for sm_100a:  tcgen05 (UMMA) fp16 candidate pass with
// TMEM accumulators and TMA-fed shared-memory tiles, followed by an exact fp32 re-rank.
//
// Replaces models/MF.py:109-112 (`user_latent @ all_item_latent.T`, cuBLAS SGEMM
// when run on a GPU), the dense score matrix + -inf mask of MF.py:117-130 and the
// per-row std::partial_sort_copy of evaluation/backend/cython/include/func.h:12-31.
//
// Exactness argument.  Tables are rescaled by powers of two and rounded to
// fp16 (unit roundoff 2^-11); the tensor cores accumulate exact products in fp32.  For a
// user u and item i,  |S~ - S| <= e(u,i) = c |u|_2 |v_i|_2,  c = 2^-10 (+5%) + d*2.4e-7.
// With L = S~ - e <= S <= S~ + e = H:  any tau with at least K CERTAINLY UNMASKED entries at
// L >= tau is a lower bound of the exact K-th score, so every item of the exact top-K has
// H >= tau.  Items are visited head | stratified sample | rest of the descending-norm order
// (reorder_kernel); a tile shares the bound e_t = c |u| * (largest norm in the tile).  The
// candidate pass keeps every item with S~ + e_t >= tau, raising tau by bisection over the
// row's candidate list when it grows (raise_fast); "certainly unmasked" comes from an exact
// bitmap of the first 64 positions and a 2048-bit Bloom filter per row.  The re-rank kernel
// drops masked items (exact test), recomputes survivors with the SAME k-ordered fp32 FMA
// chain as the exact kernel and selects by (score desc, id asc).  Rows whose list cannot be
// kept short are re-done by the exact kernel: the result always equals B200REC_SCORE_EXACT.
//
// Kernel roles (one CTA = 256 user rows = two M=128 halves, all item tiles):
//   warp 0      TMA producer: A (users) once, B (items) k-blocks of 64 fp16 through a smem ring
//   warp 1      MMA issuer: tcgen05.mma.cta_group::1.kind::f16, fp32 accumulators in TMEM,
//               tcgen05.commit frees smem slots / signals finished tiles
//                 d <= 128: M128 x N256 x K16, ONE 256-column accumulator per half, the halves
//                           ping-pong (tc_candidate_pp_kernel)
//                 d >  128: M128 x N128 x K16, two stages of a pair of 128-column accumulators
//                           (tc_candidate_kernel)
//   warps 2-9   epilogue (tc_epilogue, shared): software-pipelined tcgen05.ld 32x32b.x64
//               (thread == user row), FMNMX3 max tree against the row threshold, branch-free
//               append, register bootstrap of tau, warp-cooperative threshold raise
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>
#include <vector>
#include <cub/cub.cuh>
#include "common.cuh"
#include "topk_list.cuh"

namespace b200 {

int score_topk_exact(const float *U, const float *V, int ld, int d, const int32_t *users, int n_users, int num_items,
                     const int64_t *mi, const int32_t *mx, int k, int32_t *oi, float *os, float *dense,
                     cudaStream_t s);

constexpr int kBM = 256;        // user rows per CTA (two M=128 halves)
constexpr int kBN = 128;        // items per tile (N=128 kernel, d > 128)
constexpr int kPPN = 256;       // items per tile (N=256 ping-pong kernel, d <= 128)
constexpr int kBK = 64;         // fp16 per k-block = one 128-byte swizzle row
constexpr int kCand = 512;      // candidate slots per row
constexpr int kEpiWarps = 8;
constexpr int kThreads = (2 + kEpiWarps) * 32;
constexpr int kRowsPerLaunch = 148 * kBM * 2;

// ---- PTX wrappers ---------------------------------------------------------
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
                     "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc),
        "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float max3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start>>4 [0,14) | LBO=1 [16,30) | SBO=1024B>>4 [32,46) | version=1 [46,48) | layout=2 (SW128) [61,64)
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=B=f16 (format 0),
// K-major both, N=128 -> n_dim=16 [17,23), M=128 -> m_dim=8 [24,29)
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(kBN >> 3) << 17) | ((128u >> 4) << 24);

// ---- pre-pass kernels ----------------------------------------------------------------------------
// row norms (rounded up a hair: they feed upper bounds) + global max |x| and max norm
__global__ void __launch_bounds__(256) row_stats_kernel(const float *__restrict__ src, int ld, int d,
                                                        const int32_t *__restrict__ ids, int rows,
                                                        float *__restrict__ norms, unsigned *__restrict__ max_abs_bits) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float mabs = 0.f;
    for (int64_t r = w; r < rows; r += nw) {
        const float *p = src + (int64_t)(ids ? ids[r] : r) * ld;
        float ss = 0.f;
        for (int c = lane; c < d; c += 32) {
            const float v = p[c];
            ss = fmaf(v, v, ss);
            mabs = fmaxf(mabs, fabsf(v));
        }
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if (lane == 0) norms[r] = sqrtf(ss) * 1.000001f;
    }
    for (int o = 16; o > 0; o >>= 1) mabs = fmaxf(mabs, __shfl_xor_sync(0xffffffffu, mabs, o));
    if (lane == 0 && mabs > 0.f) atomicMax(max_abs_bits, __float_as_uint(mabs));  // nonneg floats order as uints
}

// power-of-two scale that maps max|x| into [2^13, 2^14): exact in fp32, keeps fp16 out of overflow
__device__ __forceinline__ float pow2_scale(unsigned max_abs_bits) {
    const float m = __uint_as_float(max_abs_bits);
    if (!(m > 0.f) || !isfinite(m)) return 1.f;
    int e;
    frexpf(m, &e);  // m = f * 2^e, f in [0.5, 1)
    const int sh = 14 - e;
    return ldexpf(1.f, sh > 100 ? 100 : (sh < -100 ? -100 : sh));
}

// dst[p] = fp16(scale * src[order ? order[p] : (ids ? ids[p] : p)]), rows >= `rows` and columns >= d zero
__global__ void __launch_bounds__(256) to_f16_kernel(const float *__restrict__ src, int ld, int d,
                                                     const int32_t *__restrict__ ids, const int32_t *__restrict__ order,
                                                     int rows, int rows_pad, int dpad, const unsigned *__restrict__ max_abs_bits,
                                                     __half *__restrict__ dst, float *__restrict__ scale_out) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float sc = pow2_scale(*max_abs_bits);
    if (blockIdx.x == 0 && threadIdx.x == 0 && scale_out) *scale_out = sc;
    for (int64_t r = w; r < rows_pad; r += nw) {
        const float *p = nullptr;
        if (r < rows) p = src + (int64_t)(order ? order[r] : (ids ? ids[r] : (int)r)) * ld;
        for (int c = lane; c < dpad; c += 32) dst[r * dpad + c] = __float2half_rn((p && c < d) ? p[c] * sc : 0.f);
    }
}

__global__ void inverse_perm_kernel(const int32_t *perm, int n, int32_t *inv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) inv[perm[i]] = i;
}
__global__ void iota_kernel(int32_t *out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}
// Final visiting order of the items.  Ranks are positions in the descending-norm order (one radix sort per call):
//   [0, H)            the H highest-norm items (a user aligned with the popularity direction finds its top-K here),
//   [H, H+S)          a stratified sample of the rest (every `stride`-th rank): a user whose best items sit anywhere
//                     else in the norm range (anti-aligned users: corr(score, norm) < 0) gets a usable threshold from
//                     it instead of beating its own threshold along the whole sweep,
//   [H+S, N)          everything else, still in descending-norm order.
__global__ void reorder_kernel(const int32_t *__restrict__ perm, const uint32_t *__restrict__ sorted_norm_bits, int n,
                               int H, int S, int stride, int32_t *__restrict__ perm_out, float *__restrict__ norm_out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    int pos = r;
    if (S > 0 && r >= H) {
        const int q = r - H;
        const int j = q / stride;
        if (q % stride == 0 && j < S) pos = H + j;
        else pos = H + S + q - (j + 1 < S ? j + 1 : S);
    }
    perm_out[pos] = perm[r];
    norm_out[pos] = __uint_as_float(sorted_norm_bits[r]);
}
// tile_norm[t] = largest item norm in tile t of the final order (one warp per tile)
__global__ void tile_norm_kernel(const float *__restrict__ norm, int num_items, int n_tiles, int tile, float *__restrict__ tile_norm) {
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (t >= n_tiles) return;
    float m = 0.f;
    for (int i = t * tile + lane; i < (t + 1) * tile && i < num_items; i += 32) m = fmaxf(m, norm[i]);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) tile_norm[t] = m;
}

// "Maybe masked" filters per scored row (no false negatives), one warp per row, coalesced over the row's positives:
//   wide[row][0]          EXACT bitmap of the first 64 positions (the bootstrap chunk of the ping-pong kernel)
//   wide[row][1..32]      2048-bit filter read from global memory by the (rare) threshold raises
constexpr int kWideWords = 33;   // uint64 words per row
__device__ __forceinline__ uint32_t wide_hash(uint32_t pos) { return (pos * 2654435761u) >> 21; }   // 11 bits
__global__ void __launch_bounds__(256) bloom_kernel(const int32_t *__restrict__ users, int n_rows,
                                                    const int64_t *__restrict__ mask_indptr,
                                                    const int32_t *__restrict__ mask_indices,
                                                    const int32_t *__restrict__ inv_perm,
                                                    unsigned long long *__restrict__ wide) {
    const int lane = threadIdx.x & 31;
    const int row = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (row >= n_rows) return;
    const int u = users[row];
    const int64_t mb = mask_indptr[u], me = mask_indptr[u + 1];
    unsigned long long *wrow = wide + (size_t)row * kWideWords;
    for (int q = lane; q < kWideWords; q += 32) wrow[q] = 0ull;
    __syncwarp();
    for (int64_t m = mb + lane; m < me; m += 32) {
        const uint32_t pos = (uint32_t)inv_perm[mask_indices[m]];
        if (pos < 64u) atomicOr(wrow, 1ull << pos);
        const uint32_t h2 = wide_hash(pos);
        atomicOr(wrow + 1 + (h2 >> 6), 1ull << (h2 & 63u));
    }
}

struct TcParams {
    int n_rows, num_items, n_tiles, k, d;
    const int32_t *users;            // row -> user id (mask row)
    const int64_t *mask_indptr;      // may be NULL
    const int32_t *mask_indices;
    const int32_t *inv_perm;         // item id -> sorted position
    const unsigned long long *wide;  // [n_rows][kWideWords] exact head bitmap + 2048-bit filter (NULL without a mask)
    int append_budget;               // a row that appends more than this is handed to the exact kernel
    const float *row_norm;           // [n_rows]
    const float *tile_norm;          // [n_tiles]
    const float *scale_u, *scale_v;  // device scalars (powers of two)
    uint64_t *cand;                  // [n_rows, kCand]  (ordered approx score << 32 | sorted item position)
    int32_t *cand_cnt;               // [n_rows]  (-1 = overflow: re-do with the exact kernel)
    float *dump;                     // bring-up: dense [n_rows_pad, n_tiles*kBN] approx scores (sorted order), else NULL
    int ablate;                      // diagnostics (B200REC_TC_ABLATE, ping-pong kernel): 1 no appends, 2 no filter, 4 no TMEM loads, 8 no TMA, 128 no bootstrap
    float *dbg_row;                  // diagnostics: per row [appends, final count, tau, cu]; or NULL
    unsigned long long *dbg_warp;    // diagnostics: per (block, epilogue warp) [total, wait, raise cycles, raises]; or NULL
    unsigned long long *dbg;         // diagnostics: [0] appends, [1] raises, [4] wait / [5] raise / [7] total cycles (sums), [8..11] maxima; or NULL
};

// asynchronous 64-column load (no wait) and the wait that also pins the destination registers: the
// "+r" operands make every later use of r[] depend on the wait, and keep r[] allocated in between
__device__ __forceinline__ void tmem_ld64_async(uint32_t taddr, uint32_t (&r)[64]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
        "%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,"
        "%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]),
          "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
          "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]),
          "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]),
          "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait64(uint32_t (&r)[64]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]),
                   "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]),
                   "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
    asm volatile(""
                 : "+r"(r[32]), "+r"(r[33]), "+r"(r[34]), "+r"(r[35]), "+r"(r[36]), "+r"(r[37]), "+r"(r[38]), "+r"(r[39]),
                   "+r"(r[40]), "+r"(r[41]), "+r"(r[42]), "+r"(r[43]), "+r"(r[44]), "+r"(r[45]), "+r"(r[46]), "+r"(r[47]),
                   "+r"(r[48]), "+r"(r[49]), "+r"(r[50]), "+r"(r[51]), "+r"(r[52]), "+r"(r[53]), "+r"(r[54]), "+r"(r[55]),
                   "+r"(r[56]), "+r"(r[57]), "+r"(r[58]), "+r"(r[59]), "+r"(r[60]), "+r"(r[61]), "+r"(r[62]), "+r"(r[63])
                 :
                 : "memory");
}

// Threshold raise (one row at a time, warp-cooperative, entries held in registers: one round trip to L2).
// Entries whose upper bound still reaches the new tau are compacted in place.
template <int TILE, int WM>
__device__ __forceinline__ void raise_fast(unsigned need, uint64_t *my_cand, int &cnt, float &tau, int &my_raise_at,
                                           int keff, float cu, const float *__restrict__ tile_norm, int lane,
                                           const unsigned long long *my_wide) {
    constexpr int E = kCand / 32;
    while (need) {
        const int Lsrc = __ffs(need) - 1;
        need &= need - 1;
        __syncwarp();
        uint64_t *base = reinterpret_cast<uint64_t *>(__shfl_sync(0xffffffffu, (unsigned long long)my_cand, Lsrc));
        const int n = __shfl_sync(0xffffffffu, cnt, Lsrc);
        const int kf = __shfl_sync(0xffffffffu, keff, Lsrc);
        const float c_u = __shfl_sync(0xffffffffu, cu, Lsrc);
        const float old_tau = __shfl_sync(0xffffffffu, tau, Lsrc);
        const unsigned long long *wf =
            reinterpret_cast<const unsigned long long *>(__shfl_sync(0xffffffffu, (unsigned long long)my_wide, Lsrc));
        uint64_t e[E];
        float hi[E], lo_v[E];   // upper bound of every entry; lower bound of the CLEAN (certainly unmasked) ones
#pragma unroll
        for (int i = 0; i < E; ++i) e[i] = (lane + 32 * i < n) ? base[lane + 32 * i] : 0ull;   // one round trip
        float vmin = INFINITY, vmax = -INFINITY;
        int n_clean = 0;
#pragma unroll
        for (int i = 0; i < E; ++i) {
            hi[i] = -INFINITY; lo_v[i] = -INFINITY;
            if (lane + 32 * i < n) {
                const float sc = ord2f((uint32_t)(e[i] >> 32));
                const float err = c_u * tile_norm[(uint32_t)(e[i] & 0x7FFFFFFFu) / TILE];
                hi[i] = sc + err;
                // "maybe masked" bit, computed on first sight (appends leave it clear) and kept in the entry
                if (wf) {
                    const uint32_t h2 = wide_hash((uint32_t)e[i] & 0x7FFFFFFFu);
                    e[i] |= (uint64_t)((uint32_t)(wf[1 + (h2 >> 6)] >> (h2 & 63u)) & 1u) << 31;
                }
                if (!((uint32_t)e[i] >> 31)) {
                    lo_v[i] = sc - err;
                    vmin = fminf(vmin, lo_v[i]); vmax = fmaxf(vmax, lo_v[i]);
                    ++n_clean;
                }
            }
        }
        // Selection by bisection on the value: any t with at least K clean entries at or above it is a valid
        // lower bound of the exact K-th score, so 12 halvings of [min, max] of the clean lower bounds (each a
        // per-lane count + one warp reduction) give the K-th largest to 1/4096 of the range, for any K.
        for (int o = 16; o > 0; o >>= 1) {
            vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
            vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        }
        n_clean = __reduce_add_sync(0xffffffffu, n_clean);
        float t_new = old_tau;
        if (n_clean >= kf) {
            float a = vmin, b = vmax;          // invariant: count(lo >= a) >= K
            if (kf == 1) a = vmax;
            else
                for (int it = 0; it < 12; ++it) {
                    const float mid = 0.5f * a + 0.5f * b;
                    int c = 0;
#pragma unroll
                    for (int i = 0; i < E; ++i) c += (lo_v[i] >= mid) ? 1 : 0;
                    c = __reduce_add_sync(0xffffffffu, c);
                    if (c >= kf) a = mid; else b = mid;
                }
            t_new = fmaxf(old_tau, a);
        }
        // every ballot depends on every lane's loads, so all loads have landed before the first store below
        unsigned kb[E];
#pragma unroll
        for (int i = 0; i < E; ++i) kb[i] = __ballot_sync(0xffffffffu, hi[i] >= t_new);
        int total = 0;
#pragma unroll
        for (int i = 0; i < E; ++i) {
            if ((kb[i] >> lane) & 1u) base[total + __popc(kb[i] & ((1u << lane) - 1u))] = e[i];
            total += __popc(kb[i]);
        }
        __syncwarp();
        if (lane == Lsrc) {
            // a list that stays long (many maybe-masked entries of a heavy user cannot be dropped) gets a later
            // trigger instead of a raise on every tile; past the hard limit the exact kernel re-does the row
            if (total > kCand - WM - 48) { cnt = -1; tau = INFINITY; }
            else {
                cnt = total; tau = t_new;
                if (total + 48 > my_raise_at) my_raise_at = min(kCand - WM - 1, total + 48);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Epilogue shared by both candidate kernels: thread == user row (warps 2-5 own M half 0, warps 6-9 half 1).
// PP = true : N=256 ping-pong kernel - one TILE-column accumulator per half, barriers indexed by the half.
// PP = false: N=128 kernel - two stages of a pair of TILE-column accumulators, barriers indexed by the stage.
// TMEM loads are software-pipelined (two 64-column register buffers): the load of chunk c+1 is in flight while
// chunk c is filtered, and the accumulator is handed back to the MMA warp as soon as the last load has landed -
// the filtering of the last chunk and any threshold raise run while the tensor core already works on the next tile.
// ---------------------------------------------------------------------------------------------------
template <int TILE, bool PP, bool DUMP, bool DIAG>
__device__ __forceinline__ void tc_epilogue(const TcParams &p, const int row0, const int n_tiles, const int warp,
                                            const int lane, const uint32_t tmem_base, uint64_t *tfull, uint64_t *tempty) {
    const int ew = warp - 2;
    const int q = warp & 3;
    const int h = ew >> 2;            // M half
    const int r_local = h * 128 + q * 32 + lane;
    const int row = row0 + r_local;
    const bool row_ok = row < p.n_rows;
    uint64_t *my_cand = p.cand + (size_t)(row_ok ? row : 0) * kCand;
    const float *__restrict__ tile_norm = p.tile_norm;
    const uint32_t num_items = (uint32_t)p.num_items;
    int cnt = 0, napp = 0;
    const int keff = p.k;
    const int budget = p.append_budget;
    int raise_at = min(kCand - TILE - 1, max(64, (5 * keff) / 2));   // per row: moves up when the list stays long
    float tau = -INFINITY, cu = 0.f;
    const unsigned long long *my_wide = (p.wide && row_ok) ? p.wide + (size_t)row * kWideWords : nullptr;
    if (row_ok) {
        const float c = 0.0009765625f * 1.05f + (float)p.d * 2.4e-7f;
        cu = c * p.row_norm[row] * (*p.scale_u) * (*p.scale_v);
    } else {
        tau = INFINITY;
    }
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + h * TILE;   // + stage offset (N=128 kernel)
    const int abl = DIAG ? p.ablate : 0;
    unsigned long long d_app = 0, d_raise = 0;
    long long d_wait = 0, d_rcyc = 0;
    const long long c_start = DIAG ? clock64() : 0;
    uint32_t ra[64], rb[64];
    // one 64-column chunk: FMNMX3 max tree against the row threshold.  Survivors of a hit group are appended
    // branch-free: every element is stored at the current slot and the slot only advances for a survivor (the
    // "maybe masked" bit is filled in lazily by the raise; the re-rank does the exact mask test anyway).
    auto filter_chunk = [&](const uint32_t (&r)[64], const float thr, const uint32_t pos0) {
        if (DUMP) {
            float *dst = p.dump + (size_t)(row0 + r_local) * ((size_t)n_tiles * TILE) + pos0;
#pragma unroll
            for (int j = 0; j < 64; ++j) dst[j] = __uint_as_float(r[j]);
        }
        if (abl & 2) return;
        float gm[8];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const float a = max3(__uint_as_float(r[8 * g]), __uint_as_float(r[8 * g + 1]), __uint_as_float(r[8 * g + 2]));
            const float b = max3(__uint_as_float(r[8 * g + 3]), __uint_as_float(r[8 * g + 4]), __uint_as_float(r[8 * g + 5]));
            gm[g] = max3(a, b, fmaxf(__uint_as_float(r[8 * g + 6]), __uint_as_float(r[8 * g + 7])));
        }
        const float m = max3(max3(gm[0], gm[1], gm[2]), max3(gm[3], gm[4], gm[5]), fmaxf(gm[6], gm[7]));
        if (m >= thr && cnt >= 0 && !(abl & 1)) {
            const int c_in = cnt;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                if (gm[g] >= thr) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float sc = __uint_as_float(r[8 * g + j]);
                        const uint32_t pos = pos0 + 8 * g + j;
                        my_cand[cnt] = ((uint64_t)f2ord(sc) << 32) | pos;
                        cnt += (sc >= thr && pos < num_items) ? 1 : 0;
                    }
                }
            }
            napp += cnt - c_in;
        }
    };
    for (int t = 0; t < n_tiles; ++t) {
        float thr = tau - cu * tile_norm[t];   // tau only moves at the end of a tile (and in the bootstrap)
        const long long c_w0 = (DIAG && p.dbg) ? clock64() : 0;
        const uint32_t bsel = PP ? (uint32_t)h : ((uint32_t)t & 1u);           // barrier / accumulator stage of tile t
        const uint32_t bpar = PP ? ((uint32_t)t & 1u) : (((uint32_t)t >> 1) & 1u);
        const uint32_t ta = PP ? lane_addr : lane_addr + bsel * 2 * TILE;
        mbar_wait(s32(tfull + bsel), bpar);
        if (DIAG && p.dbg) d_wait += clock64() - c_w0;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (abl & 4) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(s32(tempty + bsel));
            continue;
        }
        const uint32_t n0 = (uint32_t)t * TILE;
        {
        tmem_ld64_async(ta, ra);
        tmem_ld_wait64(ra);
        if (t == 0 && keff <= 32 && !(abl & 128)) {
            // Bootstrap (all 32 rows of the warp at once, in registers): tau0 = K-th largest lower bound among
            // the certainly-unmasked items of the first chunk - the 64 items of largest norm.  Without it every
            // row would append all of tile 0 and need a warp-cooperative raise at the same moment.
            const float e0 = cu * tile_norm[0];
            const unsigned long long head = my_wide ? my_wide[0] : 0ull;   // exact mask bitmap of positions 0..63
#pragma unroll
            for (int j = 0; j < 64; ++j)
                rb[j] = (((head >> j) & 1ull) || (uint32_t)j >= num_items) ? 0xFF800000u : ra[j];
            float prev = INFINITY;
            for (int r = 0; r < keff; ++r) {   // r-th largest distinct value (duplicates only lower the bound)
                float m = -INFINITY;
#pragma unroll
                for (int j = 0; j < 64; ++j) {
                    const float w = __uint_as_float(rb[j]);
                    m = fmaxf(m, w < prev ? w : -INFINITY);
                }
                prev = m;
            }
            if (row_ok && prev > -INFINITY) { tau = prev - e0; thr = tau - e0; }
        }
#pragma unroll
        for (int c = 0; c < TILE / 64; c += 2) {                 // ra holds chunk c (landed)
            tmem_ld64_async(ta + 64 * (c + 1), rb);
            filter_chunk(ra, thr, n0 + 64 * c);
            tmem_ld_wait64(rb);
            if (c + 2 < TILE / 64) {
                tmem_ld64_async(ta + 64 * (c + 2), ra);
                filter_chunk(rb, thr, n0 + 64 * (c + 1));
                tmem_ld_wait64(ra);
            }
        }
        }
        // every TMEM read of this warp for tile t has landed: hand the accumulator back
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(s32(tempty + bsel));
        filter_chunk(rb, thr, n0 + TILE - 64);
        // a full tile (256 appends) must always fit, and a stale tau costs appends (~100 cycles of the whole warp
        // each): raise as soon as the list holds ~2.5 K entries (>= 64), at the latest with 256 slots left
        const unsigned need = __ballot_sync(0xffffffffu, cnt > raise_at);
        if (need) {
            if (DIAG) d_raise += __popc(need);
            const long long c_r0 = (DIAG && p.dbg) ? clock64() : 0;
            raise_fast<TILE, TILE>(need, my_cand, cnt, tau, raise_at, keff, cu, tile_norm, lane, my_wide);
            if (DIAG && p.dbg) d_rcyc += clock64() - c_r0;
        }
        // a row that keeps beating its own threshold (scores rising along the sweep: a user anti-aligned with
        // the popularity direction) would drag its warp through the append path on every chunk: hand it to the
        // exact kernel instead
        if (napp > budget && cnt >= 0) { cnt = -1; tau = INFINITY; }
    }
    if (DIAG) d_app = (unsigned long long)napp;
    if (row_ok) p.cand_cnt[row] = cnt;
    if (DIAG && p.dbg_row && row_ok) {
        float *w = p.dbg_row + (size_t)row * 4;
        w[0] = (float)d_app; w[1] = (float)cnt; w[2] = tau; w[3] = cu;
    }
    if (DIAG && p.dbg) {
        atomicAdd(p.dbg + 0, d_app);
        if (lane == 0) {
            atomicAdd(p.dbg + 1, d_raise);
            atomicAdd(p.dbg + 4, (unsigned long long)d_wait); atomicAdd(p.dbg + 5, (unsigned long long)d_rcyc);
            atomicAdd(p.dbg + 7, (unsigned long long)(clock64() - c_start));
            if (p.dbg_warp) {
                unsigned long long *w = p.dbg_warp + ((size_t)blockIdx.x * kEpiWarps + ew) * 4;
                w[0] = (unsigned long long)(clock64() - c_start); w[1] = (unsigned long long)d_wait;
                w[2] = (unsigned long long)d_rcyc; w[3] = d_raise;
            }
            atomicMax(p.dbg + 8, (unsigned long long)(clock64() - c_start));
            atomicMax(p.dbg + 9, (unsigned long long)d_rcyc);
            atomicMax(p.dbg + 10, (unsigned long long)d_wait);
            atomicMax(p.dbg + 11, (unsigned long long)d_raise);
        }
    }
}

template <int KB, bool DUMP>
__global__ void __launch_bounds__(kThreads, 1)
tc_candidate_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const TcParams p, const int n_stages) {
    extern __shared__ unsigned char smem_dyn[];
    // SWIZZLE_128B operands need 1024-byte alignment: round the dynamic base up (1 KB of slack is allocated)
    unsigned char *smem_raw = smem_dyn + ((1024u - (s32(smem_dyn) & 1023u)) & 1023u);
    // layout: A[KB][256 rows][128 B] | B ring [n_stages][128 rows][128 B] | barriers | hist
    unsigned char *smA = smem_raw;
    unsigned char *smB = smA + (size_t)KB * kBM * 128;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smB + (size_t)n_stages * kBN * 128);
    uint64_t *full = bars, *empty = bars + 8, *tfull = bars + 16, *tempty = bars + 18, *afull = bars + 20;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 21);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * kBM;
    const int n_tiles = p.n_tiles;

    if (threadIdx.x == 0) {
        for (int s = 0; s < n_stages; ++s) { mbar_init(s32(full + s), 1); mbar_init(s32(empty + s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(s32(tfull + a), 1); mbar_init(s32(tempty + a), kEpiWarps); }
        mbar_init(s32(afull), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (warp == 1) {  // TMEM: all 512 columns (4 accumulators of 128 columns)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            mbar_expect_tx(s32(afull), (uint32_t)KB * kBM * 128);
            for (int kb = 0; kb < KB; ++kb) tma_load_2d(s32(smA + (size_t)kb * kBM * 128), &tmA, kb * kBK, row0, s32(afull));
            uint32_t it = 0;
            for (int t = 0; t < n_tiles; ++t) {
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const uint32_t s = it % n_stages, ph = (it / n_stages) & 1u;
                    mbar_wait(s32(empty + s), ph ^ 1u);
                    mbar_expect_tx(s32(full + s), kBN * 128);
                    tma_load_2d(s32(smB + (size_t)s * kBN * 128), &tmB, kb * kBK, t * kBN, s32(full + s));
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            mbar_wait(s32(afull), 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t it = 0;
            for (int t = 0; t < n_tiles; ++t) {
                const uint32_t as = t & 1, aph = (t >> 1) & 1u;
                mbar_wait(s32(tempty + as), aph ^ 1u);  // epilogue has drained this accumulator pair
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const uint32_t s = it % n_stages, ph = (it / n_stages) & 1u;
                    mbar_wait(s32(full + s), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t bdesc = smem_desc_sw128(s32(smB + (size_t)s * kBN * 128));
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint64_t adesc = smem_desc_sw128(s32(smA + (size_t)kb * kBM * 128 + (size_t)h * 128 * 128));
                        const uint32_t dcol = tmem_base + (as * 2 + h) * kBN;
#pragma unroll
                        for (int k = 0; k < kBK / 16; ++k)  // +32 B along K inside the swizzle atom = +2 in the address field
                            umma_f16(dcol, adesc + 2 * k, bdesc + 2 * k, kIdesc, (kb | k) ? 1u : 0u);
                    }
                    umma_commit(s32(empty + s));  // smem slot reusable once these MMAs have read it
                }
                umma_commit(s32(tfull + as));     // accumulators of tile t complete
            }
        }
    } else {
        // ================= epilogue: thread == user row (shared with the ping-pong kernel) =================
        tc_epilogue<kBN, false, DUMP, false>(p, row0, n_tiles, warp, lane, tmem_base, tfull, tempty);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}


// ---------------------------------------------------------------------------------------------------
// N=256 ping-pong variant (d <= 128).  The N=128 kernel above reads A (4 KB) + B (4 KB) from shared memory
// per 64-cycle MMA = 128 B/clk, the shared-memory limit, while TMA refills the ring: its tensor pipe is
// active only 42 % (ncu run 5).  Here one M128 x N256 x K16 instruction reads 12 KB per 128 cycles
// (96 B/clk).  TMEM holds ONE 256-column accumulator per M half; the MMA warp alternates halves, so while
// the tensor core fills half 1 the four epilogue warps of half 0 drain theirs (and vice versa).  A tile's
// k-blocks stay in the ring until both halves have consumed them.
// ---------------------------------------------------------------------------------------------------
constexpr uint32_t kIdescPP = (1u << 4) | ((uint32_t)(kPPN >> 3) << 17) | ((128u >> 4) << 24);

template <int KB, bool DUMP, bool DIAG>
__global__ void __launch_bounds__(kThreads, 1)
tc_candidate_pp_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const TcParams p, const int n_stages) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem_raw = smem_dyn + ((1024u - (s32(smem_dyn) & 1023u)) & 1023u);
    unsigned char *smA = smem_raw;                                   // [KB][256 rows][128 B]
    unsigned char *smB = smA + (size_t)KB * kBM * 128;               // ring [n_stages][256 items][128 B]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smB + (size_t)n_stages * kPPN * 128);
    uint64_t *full = bars, *empty = bars + 8, *tfull = bars + 16, *tempty = bars + 18, *afull = bars + 20;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 21);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * kBM;
    const int n_tiles = p.n_tiles;

    if (threadIdx.x == 0) {
        for (int s = 0; s < n_stages; ++s) { mbar_init(s32(full + s), 1); mbar_init(s32(empty + s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(s32(tfull + a), 1); mbar_init(s32(tempty + a), kEpiWarps / 2); }
        mbar_init(s32(afull), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {   // ---- TMA producer
            mbar_expect_tx(s32(afull), (uint32_t)KB * kBM * 128);
            for (int kb = 0; kb < KB; ++kb) tma_load_2d(s32(smA + (size_t)kb * kBM * 128), &tmA, kb * kBK, row0, s32(afull));
            uint32_t it = 0;
            for (int t = 0; t < n_tiles; ++t) {
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const uint32_t s = it % n_stages, ph = (it / n_stages) & 1u;
                    mbar_wait(s32(empty + s), ph ^ 1u);
                    if (DIAG && (p.ablate & 8) && it >= (uint32_t)n_stages) { mbar_arrive(s32(full + s)); continue; }
                    mbar_expect_tx(s32(full + s), kPPN * 128);
                    tma_load_2d(s32(smB + (size_t)s * kPPN * 128), &tmB, kb * kBK, t * kPPN, s32(full + s));
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {   // ---- MMA issuer: half 0 then half 1 of every tile
            mbar_wait(s32(afull), 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t it0 = 0;
            for (int t = 0; t < n_tiles; ++t, it0 += KB) {
                const uint32_t tph = (uint32_t)t & 1u;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    mbar_wait(s32(tempty + h), tph ^ 1u);   // the epilogue warps of this half drained tile t-1
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    for (int kb = 0; kb < KB; ++kb) {
                        const uint32_t it = it0 + kb;
                        const uint32_t s = it % n_stages, ph = (it / n_stages) & 1u;
                        if (h == 0) {
                            mbar_wait(s32(full + s), ph);
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        }
                        const uint64_t bdesc = smem_desc_sw128(s32(smB + (size_t)s * kPPN * 128));
                        const uint64_t adesc = smem_desc_sw128(s32(smA + (size_t)kb * kBM * 128 + (size_t)h * 128 * 128));
#pragma unroll
                        for (int k = 0; k < kBK / 16; ++k)
                            umma_f16(tmem_base + h * kPPN, adesc + 2 * k, bdesc + 2 * k, kIdescPP, (kb | k) ? 1u : 0u);
                        if (h == 1) umma_commit(s32(empty + s));   // both halves have read this k-block
                    }
                    umma_commit(s32(tfull + h));
                }
            }
        }
    } else {
        tc_epilogue<kPPN, true, DUMP, DIAG>(p, row0, n_tiles, warp, lane, tmem_base, tfull, tempty);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

// ---- exact fp32 re-rank of the candidates (same k-ordered FMA chain as score_exact.cu) ---------
__global__ void __launch_bounds__(256) rerank_kernel(const float *__restrict__ U, const float *__restrict__ V, int ld,
                                                     int d, const int32_t *__restrict__ users, int n_rows, int k,
                                                     const int64_t *__restrict__ mask_indptr,
                                                     const int32_t *__restrict__ mask_indices,
                                                     const int32_t *__restrict__ item_of_pos,
                                                     const uint64_t *__restrict__ cand,
                                                     const int32_t *__restrict__ cand_cnt, int32_t *__restrict__ out_idx,
                                                     float *__restrict__ out_score, int32_t *__restrict__ redo_rows,
                                                     int32_t *__restrict__ redo_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t *rk = reinterpret_cast<uint64_t *>(smem_raw) + (size_t)wid * k;
    float *us = reinterpret_cast<float *>(smem_raw + (size_t)8 * k * 8) + (size_t)wid * ((d + 3) & ~3);   // 16-byte aligned per warp
    for (int row = blockIdx.x * 8 + wid; row < n_rows; row += gridDim.x * 8) {
        const int n = cand_cnt[row];
        if (n < 0) {  // overflow / too many masked items: the exact kernel re-does this row
            if (lane == 0) redo_rows[atomicAdd(redo_count, 1)] = row;
            continue;
        }
        const int u = users[row];
        for (int c = lane; c < d; c += 32) us[c] = U[(int64_t)u * ld + c];
        for (int pp = lane; pp < k; pp += 32) rk[pp] = make_key(-INFINITY, 0x7FFFFFFF);
        __syncwarp();
        const int32_t *mrow = nullptr;
        int mdeg = 0;
        if (mask_indptr) { mrow = mask_indices + mask_indptr[u]; mdeg = (int)(mask_indptr[u + 1] - mask_indptr[u]); }
        for (int c0 = 0; c0 < n; c0 += 32) {
            const int c = c0 + lane;
            bool ok = c < n;
            uint64_t key = 0;
            if (ok) {
                const int item = item_of_pos[(uint32_t)(cand[(size_t)row * kCand + c] & 0x7FFFFFFFu)];
                int l = 0, r = mdeg;  // masked? (models/MF.py:130)
                while (l < r) { const int m = (l + r) >> 1; if (mrow[m] < item) l = m + 1; else r = m; }
                if (l < mdeg && mrow[l] == item) ok = false;
                if (ok) {
                    const float *pv = V + (int64_t)item * ld;
                    float acc = 0.f;
                    int kk = 0;
                    if ((ld & 3) == 0) {  // 16-byte row pieces; the FMA chain stays in ascending k (the oracle's order)
                        // 16 loads in flight per lane, then their 64 FMAs: 2 round trips to L2 per 128 columns
                        // instead of one per unrolled iteration
                        for (; kk + 64 <= d; kk += 64) {
                            float4 y[16];
#pragma unroll
                            for (int q = 0; q < 16; ++q) y[q] = __ldg(reinterpret_cast<const float4 *>(pv + kk + 4 * q));
#pragma unroll
                            for (int q = 0; q < 16; ++q) {
                                const float4 x = *reinterpret_cast<const float4 *>(us + kk + 4 * q);
                                acc = fmaf(x.x, y[q].x, acc); acc = fmaf(x.y, y[q].y, acc);
                                acc = fmaf(x.z, y[q].z, acc); acc = fmaf(x.w, y[q].w, acc);
                            }
                        }
                        for (; kk + 4 <= d; kk += 4) {
                            const float4 x = *reinterpret_cast<const float4 *>(us + kk);
                            const float4 y = __ldg(reinterpret_cast<const float4 *>(pv + kk));
                            acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc);
                            acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
                        }
                    }
                    for (; kk < d; ++kk) acc = fmaf(us[kk], pv[kk], acc);
                    key = make_key(acc, item);
                }
            }
            topk_list_offer(rk, k, key, ok, lane);
        }
        __syncwarp();
        // fewer than k unmasked candidates survived (tiny catalogue, a user who has nearly everything): the list still
        // ends in the sentinel - the exact kernel, which ranks the masked -inf items in id order, finishes this row
        if (key_id(rk[k - 1]) == 0x7FFFFFFF) {
            if (lane == 0) redo_rows[atomicAdd(redo_count, 1)] = row;
            __syncwarp();
            continue;
        }
        for (int pp = lane; pp < k; pp += 32) {
            out_idx[(int64_t)row * k + pp] = key_id(rk[pp]);
            if (out_score) out_score[(int64_t)row * k + pp] = key_score(rk[pp]);
        }
        __syncwarp();
    }
}

// Re-rank, staged variant (B200REC_RERANK=2; NOT the default: equal results on a B200, tests/test_gpu_experimental.py,
// but never faster than the default in a measurement).  The default kernel lets every lane walk its own candidate's V row straight from L2
// (ncu run 24: 30 % of the stalls sit on those loads inside the FMA chain); here the warp first copies the rows of up to
// RC candidates into shared memory with coalesced 16-byte loads that are all in flight at once, then RC lanes run the
// SAME k-ordered fp32 FMA chain out of shared memory (row stride d_pad + 4 floats: conflict-free float4 reads).
constexpr int kRerankRC = 8;
__global__ void __launch_bounds__(256) rerank_staged_kernel(const float *__restrict__ U, const float *__restrict__ V, int ld,
                                                            int d, const int32_t *__restrict__ users, int n_rows, int k,
                                                            const int64_t *__restrict__ mask_indptr,
                                                            const int32_t *__restrict__ mask_indices,
                                                            const int32_t *__restrict__ item_of_pos,
                                                            const uint64_t *__restrict__ cand,
                                                            const int32_t *__restrict__ cand_cnt, int32_t *__restrict__ out_idx,
                                                            float *__restrict__ out_score, int32_t *__restrict__ redo_rows,
                                                            int32_t *__restrict__ redo_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int dp = (d + 3) & ~3;                       // d rounded up to a float4; ld >= dp (tables are padded with zeros)
    const int vstride = dp + 4;
    uint64_t *rk = reinterpret_cast<uint64_t *>(smem_raw) + (size_t)wid * k;
    float *us = reinterpret_cast<float *>(smem_raw + (size_t)8 * k * 8) + (size_t)wid * dp;
    float *vs = reinterpret_cast<float *>(smem_raw + (size_t)8 * k * 8 + (size_t)8 * dp * 4) + (size_t)wid * kRerankRC * vstride;
    for (int row = blockIdx.x * 8 + wid; row < n_rows; row += gridDim.x * 8) {
        const int n = cand_cnt[row];
        if (n < 0) {
            if (lane == 0) redo_rows[atomicAdd(redo_count, 1)] = row;
            continue;
        }
        const int u = users[row];
        for (int c = lane * 4; c < dp; c += 128) *reinterpret_cast<float4 *>(us + c) = ld4(U + (int64_t)u * ld + c);
        for (int pp = lane; pp < k; pp += 32) rk[pp] = make_key(-INFINITY, 0x7FFFFFFF);
        __syncwarp();
        const int32_t *mrow = nullptr;
        int mdeg = 0;
        if (mask_indptr) { mrow = mask_indices + mask_indptr[u]; mdeg = (int)(mask_indptr[u + 1] - mask_indptr[u]); }
        for (int c0 = 0; c0 < n; c0 += kRerankRC) {
            int item = 0;
            bool ok = lane < kRerankRC && c0 + lane < n;
            if (ok) {
                item = item_of_pos[(uint32_t)(cand[(size_t)row * kCand + c0 + lane] & 0x7FFFFFFFu)];
                int l = 0, r = mdeg;  // masked? (models/MF.py:130)
                while (l < r) { const int m = (l + r) >> 1; if (mrow[m] < item) l = m + 1; else r = m; }
                if (l < mdeg && mrow[l] == item) ok = false;
            }
            const unsigned okm = __ballot_sync(0xffffffffu, ok);
#pragma unroll
            for (int j = 0; j < kRerankRC; ++j) {      // coalesced staging of the surviving rows, all loads independent
                const int it_j = __shfl_sync(0xffffffffu, item, j);
                if ((okm >> j) & 1u)
                    for (int c = lane * 4; c < dp; c += 128)
                        *reinterpret_cast<float4 *>(vs + j * vstride + c) = ldg4(V + (int64_t)it_j * ld + c);
            }
            __syncwarp();
            uint64_t key = 0;
            if (ok) {
                const float *pv = vs + lane * vstride;
                float acc = 0.f;
                int kk = 0;
                for (; kk + 4 <= d; kk += 4) {          // ascending k: the oracle's FMA order
                    const float4 x = *reinterpret_cast<const float4 *>(us + kk);
                    const float4 y = *reinterpret_cast<const float4 *>(pv + kk);
                    acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc);
                    acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
                }
                for (; kk < d; ++kk) acc = fmaf(us[kk], pv[kk], acc);
                key = make_key(acc, item);
            }
            topk_list_offer(rk, k, key, ok, lane);
            __syncwarp();
        }
        __syncwarp();
        // fewer than k unmasked candidates survived (tiny catalogue, a user who has nearly everything): the list still
        // ends in the sentinel - the exact kernel, which ranks the masked -inf items in id order, finishes this row
        if (key_id(rk[k - 1]) == 0x7FFFFFFF) {
            if (lane == 0) redo_rows[atomicAdd(redo_count, 1)] = row;
            __syncwarp();
            continue;
        }
        for (int pp = lane; pp < k; pp += 32) {
            out_idx[(int64_t)row * k + pp] = key_id(rk[pp]);
            if (out_score) out_score[(int64_t)row * k + pp] = key_score(rk[pp]);
        }
        __syncwarp();
    }
}

__global__ void gather_ids_kernel(const int32_t *users, const int32_t *rows, int n, int32_t *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = users[rows[i]];
}
__global__ void scatter_rows_kernel(const int32_t *rows, int n, int k, const int32_t *src_idx, const float *src_sc,
                                    int32_t *dst_idx, float *dst_sc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n * k) {
        const int r = rows[i / k], c = i % k;
        dst_idx[(int64_t)r * k + c] = src_idx[i];
        if (dst_sc) dst_sc[(int64_t)r * k + c] = src_sc[i];
    }
}

// ---- host -------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// fp16 [rows, dpad] row-major; box = 64 elements (128 B, SWIZZLE_128B) x box_rows
static int make_map(CUtensorMap *m, void *base, uint64_t rows, uint64_t dpad, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    B200_REQUIRE(fn != nullptr, B200REC_ECUDA, "cuTensorMapEncodeTiled entry point not found");
    cuuint64_t dims[2] = {dpad, rows};
    cuuint64_t strides[1] = {dpad * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_REQUIRE(r == CUDA_SUCCESS, B200REC_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return B200REC_OK;
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static size_t sort_temp_bytes(int num_items) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                              (const int32_t *)nullptr, (int32_t *)nullptr, num_items, 16, 32);
    return bytes;
}

struct TcLayout {
    int dpad, KB, rows_cap, items_pad, n_tiles, tile;
    size_t off_vh, off_uh, off_unorm, off_vnorm, off_vnorm_sorted, off_iota, off_perm, off_inv, off_tnorm, off_scalars, off_sort,
        sort_bytes, off_cand, off_cnt, off_redo, off_redo_n, off_ridx, off_rsc, off_ruser, off_wide, total;
};
static TcLayout tc_layout(int n_users, int num_items, int d, int k) {
    TcLayout L;
    L.dpad = (int)align_up(d, kBK);
    L.KB = L.dpad / kBK;
    L.rows_cap = n_users < kRowsPerLaunch ? (int)align_up(n_users > 0 ? n_users : 1, kBM) : kRowsPerLaunch;
    // d <= 128: N=256 ping-pong kernel (ncu r16: tensor pipe 91.6 % of peak sustained active, 0.58 ms for 32768 x 100k);
    // d > 128: N=128 kernel (the ping-pong ring would need 2*KB stages of 32 KB).  B200REC_TC_KERNEL=0 forces N=128.
    const char *kenv = getenv("B200REC_TC_KERNEL");
    const bool use_pp = !(kenv && atoi(kenv) == 0);
    L.tile = (use_pp && L.KB <= 2) ? kPPN : kBN;
    L.items_pad = (int)align_up(num_items, L.tile);
    L.n_tiles = L.items_pad / L.tile;
    L.sort_bytes = sort_temp_bytes(num_items);
    size_t o = 0;
    auto take = [&](size_t bytes, size_t al) { size_t at = align_up(o, al); o = at + bytes; return at; };
    L.off_vh = take((size_t)L.items_pad * L.dpad * 2, 1024);
    L.off_uh = take((size_t)L.rows_cap * L.dpad * 2, 1024);
    L.off_unorm = take((size_t)L.rows_cap * 4, 256);
    L.off_vnorm = take((size_t)L.items_pad * 4, 256);
    L.off_vnorm_sorted = take((size_t)L.items_pad * 4, 256);
    L.off_iota = take((size_t)L.items_pad * 4, 256);
    L.off_perm = take((size_t)L.items_pad * 4, 256);
    L.off_inv = take((size_t)L.items_pad * 4, 256);
    L.off_tnorm = take((size_t)L.n_tiles * 4, 256);
    L.off_scalars = take(256, 256);   // [0] max|v| bits, [1] max|u| bits, [2] scale_v, [3] scale_u
    L.off_sort = take(L.sort_bytes, 256);
    L.off_cand = take((size_t)L.rows_cap * kCand * 8, 256);
    L.off_cnt = take((size_t)L.rows_cap * 4, 256);
    L.off_redo = take((size_t)L.rows_cap * 4, 256);
    L.off_redo_n = take(4, 256);
    L.off_ruser = take((size_t)L.rows_cap * 4, 256);
    L.off_wide = take((size_t)L.rows_cap * kWideWords * 8, 256);
    L.off_ridx = take((size_t)L.rows_cap * k * 4, 256);
    L.off_rsc = take((size_t)L.rows_cap * k * 4, 256);
    L.total = align_up(o, 256);
    return L;
}

int64_t score_topk_tc_workspace(int n_users, int num_items, int d, int k) {
    return (int64_t)tc_layout(n_users, num_items, d, k).total + 1024;
}

template <int KB>
static int launch_candidates(const CUtensorMap &ma, const CUtensorMap &mb, const TcParams &p, int n_blocks,
                             cudaStream_t s) {
    const size_t a_bytes = (size_t)KB * kBM * 128;
    int stages = (int)((224 * 1024 - a_bytes - 4096) / (kBN * 128));
    if (stages > 8) stages = 8;
    B200_REQUIRE(stages >= 2, B200REC_EUNSUPPORTED, "score_topk TC: d too large for shared memory");
    const size_t smem = 1024 + a_bytes + (size_t)stages * kBN * 128 + 22 * 8 + kEpiWarps * 32 * 4 + 64;
    if (p.dump) {
        auto kern = tc_candidate_kernel<KB, true>;
        B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<n_blocks, kThreads, smem, s>>>(ma, mb, p, stages);
    } else {
        auto kern = tc_candidate_kernel<KB, false>;
        B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<n_blocks, kThreads, smem, s>>>(ma, mb, p, stages);
    }
    B200_LAUNCH_CHECK();
    return B200REC_OK;
}

template <int KB>
static int launch_candidates_pp(const CUtensorMap &ma, const CUtensorMap &mb, const TcParams &p, int n_blocks,
                                cudaStream_t s) {
    const size_t a_bytes = (size_t)KB * kBM * 128;
    int stages = (int)((224 * 1024 - a_bytes - 4096) / (kPPN * 128));
    if (stages > 8) stages = 8;
    B200_REQUIRE(stages >= 2 * KB, B200REC_EUNSUPPORTED, "score_topk TC: ping-pong kernel needs 2 tiles of stages");
    const size_t smem = 1024 + a_bytes + (size_t)stages * kPPN * 128 + 22 * 8 + kEpiWarps * 32 * 4 + 64;
    if (p.dump) {
        auto kern = tc_candidate_pp_kernel<KB, true, false>;
        B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<n_blocks, kThreads, smem, s>>>(ma, mb, p, stages);
    } else if (p.dbg || p.ablate) {
        auto kern = tc_candidate_pp_kernel<KB, false, true>;
        B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<n_blocks, kThreads, smem, s>>>(ma, mb, p, stages);
    } else {
        auto kern = tc_candidate_pp_kernel<KB, false, false>;
        B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<n_blocks, kThreads, smem, s>>>(ma, mb, p, stages);
    }
    B200_LAUNCH_CHECK();
    return B200REC_OK;
}

int score_topk_tc_impl(const float *U, const float *V, int ld, int d, const int32_t *users, int n_users,
                       int num_items, const int64_t *mi, const int32_t *mx, int k, int32_t *oi, float *os, void *ws,
                       int64_t ws_bytes, float *dump, cudaStream_t s) {
    if (n_users <= 0) return B200REC_OK;
    B200_REQUIRE(d <= 256, B200REC_EUNSUPPORTED, "score_topk TC: d <= 256 (got %d)", d);
    B200_REQUIRE(k <= kCand / 4, B200REC_EUNSUPPORTED, "score_topk TC: k <= %d (got %d)", kCand / 4, k);
    const TcLayout L = tc_layout(n_users, num_items, d, k);
    unsigned char *base = reinterpret_cast<unsigned char *>(align_up((size_t)ws, 1024));
    B200_REQUIRE(ws && (int64_t)((base - (unsigned char *)ws) + L.total) <= ws_bytes, B200REC_ENOMEM,
                 "score_topk TC: workspace too small (%lld < %lld)", (long long)ws_bytes, (long long)L.total + 1024);
    __half *vh = reinterpret_cast<__half *>(base + L.off_vh);
    __half *uh = reinterpret_cast<__half *>(base + L.off_uh);
    float *unorm = reinterpret_cast<float *>(base + L.off_unorm);
    float *vnorm = reinterpret_cast<float *>(base + L.off_vnorm);
    uint32_t *vnorm_sorted = reinterpret_cast<uint32_t *>(base + L.off_vnorm_sorted);
    int32_t *iota = reinterpret_cast<int32_t *>(base + L.off_iota);
    int32_t *perm = reinterpret_cast<int32_t *>(base + L.off_perm);
    int32_t *inv_perm = reinterpret_cast<int32_t *>(base + L.off_inv);
    float *tnorm = reinterpret_cast<float *>(base + L.off_tnorm);
    unsigned *scal = reinterpret_cast<unsigned *>(base + L.off_scalars);
    uint64_t *cand = reinterpret_cast<uint64_t *>(base + L.off_cand);
    int32_t *cnt = reinterpret_cast<int32_t *>(base + L.off_cnt);
    int32_t *redo = reinterpret_cast<int32_t *>(base + L.off_redo);
    int32_t *redo_n = reinterpret_cast<int32_t *>(base + L.off_redo_n);
    int32_t *ruser = reinterpret_cast<int32_t *>(base + L.off_ruser);
    int32_t *ridx = reinterpret_cast<int32_t *>(base + L.off_ridx);
    float *rsc = reinterpret_cast<float *>(base + L.off_rsc);
    const int sms = sm_count();

    // ---- item side, once per call: norms, descending-norm order, fp16 copy in that order, per-tile bound ----
    B200_CUDA(cudaMemsetAsync(scal, 0, 256, s));
    row_stats_kernel<<<sms * 8, 256, 0, s>>>(V, ld, d, nullptr, num_items, vnorm, scal + 0);
    B200_LAUNCH_CHECK();
    iota_kernel<<<(num_items + 255) / 256, 256, 0, s>>>(iota, num_items);
    B200_LAUNCH_CHECK();
    size_t sort_bytes = L.sort_bytes;
    B200_CUDA(cub::DeviceRadixSort::SortPairsDescending(base + L.off_sort, sort_bytes,
                                                         reinterpret_cast<const uint32_t *>(vnorm), vnorm_sorted,
                                                         (const int32_t *)iota, perm, num_items, 16, 32, s));
    count_launch(3);
    // final visiting order (head | stratified sample | rest): the iota and unsorted-norm buffers are free again
    int32_t *order = iota;
    float *norm_final = vnorm;
    {
        const int H = L.tile, S = 16 * L.tile;
        const bool sample = num_items >= 8 * (H + S);
        const int stride = sample ? (num_items - H) / S : 1;
        reorder_kernel<<<(num_items + 255) / 256, 256, 0, s>>>(perm, vnorm_sorted, num_items, H, sample ? S : 0, stride, order,
                                                                norm_final);
        B200_LAUNCH_CHECK();
    }
    inverse_perm_kernel<<<(num_items + 255) / 256, 256, 0, s>>>(order, num_items, inv_perm);
    B200_LAUNCH_CHECK();
    to_f16_kernel<<<sms * 8, 256, 0, s>>>(V, ld, d, nullptr, order, num_items, L.items_pad, L.dpad, scal + 0, vh,
                                           reinterpret_cast<float *>(scal + 2));
    B200_LAUNCH_CHECK();
    tile_norm_kernel<<<(L.n_tiles * 32 + 255) / 256, 256, 0, s>>>(norm_final, num_items, L.n_tiles, L.tile, tnorm);
    B200_LAUNCH_CHECK();
    CUtensorMap mb;
    int rc = make_map(&mb, vh, (uint64_t)L.items_pad, (uint64_t)L.dpad, (uint32_t)L.tile);
    if (rc) return rc;

    for (int r0 = 0; r0 < n_users; r0 += L.rows_cap) {
        const int nr = (n_users - r0) < L.rows_cap ? (n_users - r0) : L.rows_cap;
        const int nr_pad = (int)align_up(nr, kBM);
        B200_CUDA(cudaMemsetAsync(scal + 1, 0, 4, s));
        row_stats_kernel<<<sms * 4, 256, 0, s>>>(U, ld, d, users + r0, nr, unorm, scal + 1);
        B200_LAUNCH_CHECK();
        to_f16_kernel<<<sms * 4, 256, 0, s>>>(U, ld, d, users + r0, nullptr, nr, nr_pad, L.dpad, scal + 1, uh,
                                               reinterpret_cast<float *>(scal + 3));
        B200_LAUNCH_CHECK();
        CUtensorMap ma;
        if ((rc = make_map(&ma, uh, (uint64_t)nr_pad, (uint64_t)L.dpad, kBM))) return rc;
        TcParams p;
        p.n_rows = nr; p.num_items = num_items; p.n_tiles = L.n_tiles; p.k = k; p.d = d;
        p.users = users + r0; p.mask_indptr = mi; p.mask_indices = mx; p.inv_perm = inv_perm;
        p.wide = nullptr;
        p.append_budget = 1536 + 8 * k;
#ifdef B200REC_TC_DIAG
        if (getenv("B200REC_TC_BUDGET")) p.append_budget = atoi(getenv("B200REC_TC_BUDGET"));
#endif
        if (mi) {
            unsigned long long *wide = reinterpret_cast<unsigned long long *>(base + L.off_wide);
            bloom_kernel<<<(nr + 7) / 8, 256, 0, s>>>(users + r0, nr, mi, mx, inv_perm, wide);
            B200_LAUNCH_CHECK();
            p.wide = wide;
        }
        p.row_norm = unorm; p.tile_norm = tnorm;
        p.scale_v = reinterpret_cast<float *>(scal + 2); p.scale_u = reinterpret_cast<float *>(scal + 3);
        p.cand = cand; p.cand_cnt = cnt; p.dump = dump ? dump + (size_t)r0 * L.items_pad : nullptr;
        // diagnostics (ablation flags, kernel time and filter counters on stderr) are compiled in only with
        // -DB200REC_TC_DIAG (tools/build_ablate.sh); the product library carries none of it
        p.ablate = 0; p.dbg = nullptr; p.dbg_warp = nullptr; p.dbg_row = nullptr;
#ifdef B200REC_TC_DIAG
        const char *abl_env = getenv("B200REC_TC_ABLATE");
        const bool diag = getenv("B200REC_TC_TIME") != nullptr;
        p.ablate = abl_env ? atoi(abl_env) : 0;
        p.dbg = (diag && atoi(getenv("B200REC_TC_TIME")) >= 2) ? reinterpret_cast<unsigned long long *>(scal + 8) : nullptr;
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        p.dbg_warp = nullptr; p.dbg_row = nullptr;
        if (p.dbg && atoi(getenv("B200REC_TC_TIME")) >= 4) {
            cudaMalloc(&p.dbg_row, (size_t)nr * 16);
            cudaMemset(p.dbg_row, 0, (size_t)nr * 16);
        }
        if (p.dbg && atoi(getenv("B200REC_TC_TIME")) == 3) {
            cudaMalloc(&p.dbg_warp, (size_t)(nr_pad / kBM) * kEpiWarps * 32);
            cudaMemset(p.dbg_warp, 0, (size_t)(nr_pad / kBM) * kEpiWarps * 32);
        }
        if (diag) {
            B200_CUDA(cudaMemsetAsync(scal + 8, 0, 128, s));
            cudaEventCreate(&ev0); cudaEventCreate(&ev1);
            cudaEventRecord(ev0, s);
        }
#endif
        switch (L.KB) {
            case 1: rc = (L.tile == kPPN) ? launch_candidates_pp<1>(ma, mb, p, nr_pad / kBM, s)
                                          : launch_candidates<1>(ma, mb, p, nr_pad / kBM, s); break;
            case 2: rc = (L.tile == kPPN) ? launch_candidates_pp<2>(ma, mb, p, nr_pad / kBM, s)
                                          : launch_candidates<2>(ma, mb, p, nr_pad / kBM, s); break;
            case 3: rc = launch_candidates<3>(ma, mb, p, nr_pad / kBM, s); break;
            default: rc = launch_candidates<4>(ma, mb, p, nr_pad / kBM, s); break;
        }
        if (rc) return rc;
#ifdef B200REC_TC_DIAG
        if (diag) {
            cudaEventRecord(ev1, s);
            cudaEventSynchronize(ev1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev0, ev1);
            unsigned long long h[16];
            cudaMemcpy(h, scal + 8, 128, cudaMemcpyDeviceToHost);
            fprintf(stderr, "[b200rec tc] candidate kernel tile=%d ablate=%d rows=%d: %.3f ms  appends/row=%.1f raises/row=%.2f"
                            "\n", L.tile, p.ablate, nr, ms, (double)h[0] / nr, (double)h[1] / nr);
            if (h[7])
                fprintf(stderr, "        epilogue warp cycles: wait tfull %.1f%%  raise %.1f%%  (mean total %.0f cycles/warp)\n",
                        100.0 * h[4] / h[7], 100.0 * h[5] / h[7],
                        (double)h[7] / ((double)(nr_pad / kBM) * kEpiWarps));
            if (h[7])
                fprintf(stderr, "        max over warps: total %llu cycles, raise %llu cycles, wait %llu cycles, raises %llu\n", h[8], h[9],
                        h[10], h[11]);
            if (p.dbg_warp) {
                const int nb = nr_pad / kBM;
                std::vector<unsigned long long> hw((size_t)nb * kEpiWarps * 4);
                cudaMemcpy(hw.data(), p.dbg_warp, hw.size() * 8, cudaMemcpyDeviceToHost);
                for (int b = 0; b < nb; ++b) {
                    unsigned long long mx = 0;
                    for (int w = 0; w < kEpiWarps; ++w) mx = hw[((size_t)b * kEpiWarps + w) * 4] > mx ? hw[((size_t)b * kEpiWarps + w) * 4] : mx;
                    if (b < 4 || mx > 1800000ull) {
                        fprintf(stderr, "        cta %3d:", b);
                        for (int w = 0; w < kEpiWarps; ++w) {
                            const unsigned long long *q = &hw[((size_t)b * kEpiWarps + w) * 4];
                            fprintf(stderr, " [%lluk w%lluk r%lluk n%llu]", q[0] / 1000, q[1] / 1000, q[2] / 1000, q[3]);
                        }
                        fprintf(stderr, "\n");
                    }
                }
                cudaFree(p.dbg_warp);
            }
            if (p.dbg_row) {
                std::vector<float> hr((size_t)nr * 4), hn((size_t)nr), tn((size_t)L.n_tiles);
                cudaMemcpy(hr.data(), p.dbg_row, hr.size() * 4, cudaMemcpyDeviceToHost);
                cudaMemcpy(hn.data(), unorm, (size_t)nr * 4, cudaMemcpyDeviceToHost);
                cudaMemcpy(tn.data(), tnorm, (size_t)L.n_tiles * 4, cudaMemcpyDeviceToHost);
                float sc[2];
                cudaMemcpy(sc, scal + 2, 8, cudaMemcpyDeviceToHost);
                fprintf(stderr, "        scale_v=%g scale_u=%g tile_norm[0]=%g [1]=%g [last]=%g\n", sc[0], sc[1], tn[0], tn[1], tn[L.n_tiles - 1]);
                int shown = 0;
                for (int r = 0; r < nr && shown < 24; ++r)
                    if (hr[(size_t)r * 4] > 1000.f) {
                        fprintf(stderr, "        row %5d: appends %.0f cnt %.0f tau %g cu %g |u| %g\n", r, hr[(size_t)r * 4],
                                hr[(size_t)r * 4 + 1], hr[(size_t)r * 4 + 2], hr[(size_t)r * 4 + 3], hn[r]);
                        ++shown;
                    }
                cudaFree(p.dbg_row);
            }
            cudaEventDestroy(ev0); cudaEventDestroy(ev1);
        }
#endif
        if (!oi) continue;  // dump-only bring-up call
        B200_CUDA(cudaMemsetAsync(redo_n, 0, 4, s));
        const size_t rsmem = (size_t)8 * k * 8 + (size_t)8 * ((d + 3) & ~3) * 4;
        B200_CUDA(cudaFuncSetAttribute(rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem));
        int rgrid = (nr + 7) / 8;
        {   // whole waves of resident CTAs (registers decide how many fit), rows are grid-strided
            int occ = 0;
            B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rerank_kernel, 256, rsmem));
            if (occ < 1) occ = 1;
            if (rgrid > sms * occ) rgrid = sms * occ;
        }
#ifdef B200REC_TC_DIAG
        cudaEvent_t er0 = nullptr, er1 = nullptr;
        if (diag) { cudaEventCreate(&er0); cudaEventCreate(&er1); cudaEventRecord(er0, s); }
#endif
        const bool staged = getenv("B200REC_RERANK") && atoi(getenv("B200REC_RERANK")) == 2 && (ld & 3) == 0;
        if (staged) {   // experimental (see rerank_staged_kernel); the default stays the validated kernel
            const size_t dp = (size_t)((d + 3) & ~3);
            const size_t ssmem = (size_t)8 * k * 8 + 8 * dp * 4 + (size_t)8 * kRerankRC * (dp + 4) * 4;
            B200_CUDA(cudaFuncSetAttribute(rerank_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssmem));
            int occ = 0, sgrid = (nr + 7) / 8;
            B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rerank_staged_kernel, 256, ssmem));
            if (occ < 1) occ = 1;
            if (sgrid > sms * occ) sgrid = sms * occ;
            rerank_staged_kernel<<<sgrid, 256, ssmem, s>>>(U, V, ld, d, users + r0, nr, k, mi, mx, order, cand, cnt,
                                                          oi + (size_t)r0 * k, os ? os + (size_t)r0 * k : nullptr, redo, redo_n);
        } else {
            rerank_kernel<<<rgrid, 256, rsmem, s>>>(U, V, ld, d, users + r0, nr, k, mi, mx, order, cand, cnt,
                                                    oi + (size_t)r0 * k, os ? os + (size_t)r0 * k : nullptr, redo, redo_n);
        }
        B200_LAUNCH_CHECK();
#ifdef B200REC_TC_DIAG
        if (diag) {
            cudaEventRecord(er1, s); cudaEventSynchronize(er1);
            float ms = 0.f; cudaEventElapsedTime(&ms, er0, er1);
            fprintf(stderr, "[b200rec tc] rerank kernel rows=%d: %.3f ms\n", nr, ms);
            cudaEventDestroy(er0); cudaEventDestroy(er1);
        }
#endif
        int n_redo = 0;
        B200_CUDA(cudaMemcpyAsync(&n_redo, redo_n, 4, cudaMemcpyDeviceToHost, s));
        B200_CUDA(cudaStreamSynchronize(s));
#ifdef B200REC_TC_DIAG
        if (getenv("B200REC_TC_STATS")) {  // diagnostics only
            std::vector<int32_t> hc((size_t)nr);
            cudaMemcpy(hc.data(), cnt, (size_t)nr * 4, cudaMemcpyDeviceToHost);
            long long tot = 0, mxc = 0, over = 0;
            for (int v : hc) { if (v < 0) ++over; else { tot += v; if (v > mxc) mxc = v; } }
            fprintf(stderr, "[b200rec tc] rows=%d redo=%d (cnt<0: %lld) mean_cand=%.1f max_cand=%lld k=%d tiles=%d\n", nr,
                    n_redo, over, nr > over ? (double)tot / (double)(nr - over) : 0.0, mxc, k, p.n_tiles);
        }
#endif
        if (n_redo > 0) {  // rows the candidate buffer could not hold: exact kernel, then scatter back
            gather_ids_kernel<<<(n_redo + 255) / 256, 256, 0, s>>>(users + r0, redo, n_redo, ruser);
            B200_LAUNCH_CHECK();
            if ((rc = score_topk_exact(U, V, ld, d, ruser, n_redo, num_items, mi, mx, k, ridx, os ? rsc : nullptr,
                                       nullptr, s)))
                return rc;
            scatter_rows_kernel<<<(n_redo * k + 255) / 256, 256, 0, s>>>(redo, n_redo, k, ridx, os ? rsc : nullptr,
                                                                         oi + (size_t)r0 * k,
                                                                         os ? os + (size_t)r0 * k : nullptr);
            B200_LAUNCH_CHECK();
        }
    }
    return B200REC_OK;
}

int score_topk_tc(const float *U, const float *V, int ld, int d, const int32_t *users, int n_users, int num_items,
                  const int64_t *mi, const int32_t *mx, int k, int32_t *oi, float *os, void *ws, int64_t ws_bytes,
                  cudaStream_t s) {
    return score_topk_tc_impl(U, V, ld, d, users, n_users, num_items, mi, mx, k, oi, os, ws, ws_bytes, nullptr, s);
}

}  // namespace b200

// bring-up / test hook: raw fp16 tensor-core scores of the candidate pass, dense fp32
// [rows_pad(256), items_pad(128)] row-major, items in the kernel's descending-norm order and
// in the rescaled domain; `perm_out` (int32 [num_items]) and `scales_out` (2 floats: scale_v,
// scale_u) describe both.  Not part of the reference surface.
extern "C" int b200rec_debug_tc_scores(const float *U, const float *V, int ld, int d, const int32_t *users, int n_users,
                                       int num_items, float *dump, void *workspace, int64_t workspace_bytes,
                                       void *stream) {
    using namespace b200;
    B200_REQUIRE(U && V && users && dump && workspace, B200REC_EINVAL, "debug_tc_scores: null argument");
    B200_REQUIRE(n_users <= kRowsPerLaunch, B200REC_EINVAL, "debug_tc_scores: at most %d rows", kRowsPerLaunch);
    return score_topk_tc_impl(U, V, ld, d, users, n_users, num_items, nullptr, nullptr, 1, nullptr, nullptr, workspace,
                              workspace_bytes, dump, (cudaStream_t)stream);
}

// workspace offsets of the permutation and the two scales written by the call above (test hook)
extern "C" int b200rec_debug_tc_layout(int n_users, int num_items, int d, int64_t *off_perm, int64_t *off_scales) {
    using namespace b200;
    const TcLayout L = tc_layout(n_users, num_items, d, 1);
    *off_perm = (int64_t)L.off_iota;   // the final visiting order lives in the iota buffer
    *off_scales = (int64_t)L.off_scalars + 8;
    return B200REC_OK;
}
