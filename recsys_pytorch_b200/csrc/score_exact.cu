// Exact fp32 scoring + masked top-K on CUDA cores (sm_100a).
//
// Replaces models/MF.py:109-132 (predict_batch_users GEMM, D2H into a dense
// float64 [U,I] matrix, -inf mask) and evaluation/backend/cython/include/
// func.h:12-31 (per-row std::partial_sort_copy) with one kernel that never
// materialises the score matrix: per 64-user block, item tiles are scored with a
// k-ordered fp32 FMA chain (bit-identical to oracle/eval_oracle.c::
// oracle_score_topk_chunk), masked from the CSR row, and merged into a per-row
// sorted top-K list in shared memory keyed (score desc, id asc).
//
// This is the exactness anchor of the scoring path: the tcgen05 kernel
// (score_tc.cu) uses the same list/merge code for its fp32 re-rank and falls back
// to this kernel for rows whose candidate set overflows.
#include <math.h>
#include "common.cuh"
#include "topk_list.cuh"

namespace b200 {

constexpr int kKSlab = 32;  // k elements staged per shared-memory slab
constexpr int kPad = kKSlab + 4;  // row stride of a staged slab (bank-conflict-free float4 access)

// TM users x TN items per tile, 256 threads, each thread a 4x4 register tile.
template <int TM>
struct ExactCfg {
    static constexpr int TN = 4096 / TM;
    static constexpr int TR = TM / 4;   // thread rows
    static constexpr int TC = 256 / TR; // thread cols; TC*4 == TN
};

template <int TM>
__global__ void __launch_bounds__(256) score_topk_exact_kernel(
    const float *__restrict__ U, const float *__restrict__ V, int ld, int d, const int32_t *__restrict__ users,
    int n_users, int num_items, const int64_t *__restrict__ mask_indptr, const int32_t *__restrict__ mask_indices,
    int k, int32_t *__restrict__ out_idx, float *__restrict__ out_score, float *__restrict__ dense_out,
    int items_per_split, uint64_t *__restrict__ part_keys) {
    // blockIdx.y = item split: with few user rows the catalogue is divided over gridDim.y CTAs per row
    // block (each writes its local top-k keys to part_keys[row][split][k]; merge_topk_kernel finishes)
    const int item_begin = blockIdx.y * items_per_split;
    const int item_end = (item_begin + items_per_split < num_items) ? item_begin + items_per_split : num_items;
    using C = ExactCfg<TM>;
    constexpr int TN = C::TN;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: Us[TM][kPad] | Vs[TN][kPad] | S[TM][TN+1] | mcur[TM] (int64) | keys[TM][k] (u64)
    float *Us = reinterpret_cast<float *>(smem_raw);
    float *Vs = Us + TM * kPad;
    float *S = Vs + TN * kPad;
    int64_t *mcur = reinterpret_cast<int64_t *>(S + TM * (TN + 1) + ((TM * (TN + 1)) & 1));
    uint64_t *keys = reinterpret_cast<uint64_t *>(mcur + TM);

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int row0 = blockIdx.x * TM;
    const int tr = tid / C::TC, tc = tid % C::TC;  // thread tile: rows 4*tr+a, cols tc+TC*b

    // per-row state
    for (int r = tid; r < TM; r += 256) {
        const int gr = row0 + r;
        mcur[r] = (mask_indptr && gr < n_users) ? mask_indptr[users[gr]] : 0;
    }
    if (k > 0)
        for (int e = tid; e < TM * k; e += 256) keys[e] = make_key(-INFINITY, 0x7FFFFFFF);
    __syncthreads();

    const bool vec_ok = (ld % 4) == 0;
    for (int n0 = item_begin; n0 < item_end; n0 += TN) {
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

        for (int k0 = 0; k0 < d; k0 += kKSlab) {
            // stage row-major slabs (k contiguous, row stride kPad): Us[r][kk], Vs[c][kk]; columns >= d read as 0
            for (int e = tid; e < TM * (kKSlab / 4); e += 256) {
                const int r = e / (kKSlab / 4), kq = e % (kKSlab / 4);
                const int gr = row0 + r, kk = k0 + 4 * kq;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (gr < n_users && kk < d) {
                    const float *src = U + (int64_t)users[gr] * ld + kk;
                    if (vec_ok && kk + 3 < ld) v = *reinterpret_cast<const float4 *>(src);
                    else { v.x = src[0]; if (kk + 1 < ld) v.y = src[1]; if (kk + 2 < ld) v.z = src[2]; if (kk + 3 < ld) v.w = src[3]; }
                    if (kk + 1 >= d) v.y = 0.f; if (kk + 2 >= d) v.z = 0.f; if (kk + 3 >= d) v.w = 0.f;
                }
                *reinterpret_cast<float4 *>(Us + r * kPad + 4 * kq) = v;
            }
            for (int e = tid; e < TN * (kKSlab / 4); e += 256) {
                const int c = e / (kKSlab / 4), kq = e % (kKSlab / 4);
                const int kk = k0 + 4 * kq;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (n0 + c < item_end && kk < d) {
                    const float *src = V + (int64_t)(n0 + c) * ld + kk;
                    if (vec_ok && kk + 3 < ld) v = *reinterpret_cast<const float4 *>(src);
                    else { v.x = src[0]; if (kk + 1 < ld) v.y = src[1]; if (kk + 2 < ld) v.z = src[2]; if (kk + 3 < ld) v.w = src[3]; }
                    if (kk + 1 >= d) v.y = 0.f; if (kk + 2 >= d) v.z = 0.f; if (kk + 3 >= d) v.w = 0.f;
                }
                *reinterpret_cast<float4 *>(Vs + c * kPad + 4 * kq) = v;
            }
            __syncthreads();
#pragma unroll 2
            for (int kk = 0; kk < kKSlab; kk += 4) {  // ascending k: the oracle's FMA order (zeros beyond d)
                float4 u4[4], v4[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) u4[a] = *reinterpret_cast<const float4 *>(Us + (4 * tr + a) * kPad + kk);
#pragma unroll
                for (int b = 0; b < 4; ++b) v4[b] = *reinterpret_cast<const float4 *>(Vs + (tc + C::TC * b) * kPad + kk);
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        acc[a][b] = fmaf(u4[a].x, v4[b].x, acc[a][b]);
                        acc[a][b] = fmaf(u4[a].y, v4[b].y, acc[a][b]);
                        acc[a][b] = fmaf(u4[a].z, v4[b].z, acc[a][b]);
                        acc[a][b] = fmaf(u4[a].w, v4[b].w, acc[a][b]);
                    }
            }
            __syncthreads();
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) S[(4 * tr + a) * (TN + 1) + tc + C::TC * b] = acc[a][b];
        __syncthreads();

        // warp-per-row: mask (models/MF.py:130), optional dense write, top-K merge (func.h:12-20)
        for (int r = wid; r < TM; r += 8) {
            const int gr = row0 + r;
            if (gr >= n_users) continue;
            float *Srow = S + r * (TN + 1);
            if (mask_indptr) {
                const int64_t end = mask_indptr[users[gr] + 1];
                int64_t cur = mcur[r];
                for (;;) {
                    const int64_t p = cur + lane;
                    int it = (p < end) ? mask_indices[p] : 0x7FFFFFFF;
                    const bool in = it < n0 + TN;
                    if (in && it >= n0) Srow[it - n0] = -INFINITY;
                    const unsigned b = __ballot_sync(0xffffffffu, in);
                    cur += __popc(b);
                    if (b != 0xffffffffu) break;
                }
                __syncwarp();
                if (lane == 0) mcur[r] = cur;
            }
            if (dense_out) {
                for (int c = lane; c < TN && n0 + c < item_end; c += 32)
                    dense_out[(int64_t)gr * num_items + n0 + c] = Srow[c];
            }
            if (k > 0) {
                uint64_t *rk = keys + (size_t)r * k;
                for (int c0 = 0; c0 < TN; c0 += 32) {
                    const int c = c0 + lane;
                    const bool ok = (c < TN) && (n0 + c < item_end);
                    const uint64_t key = ok ? make_key(Srow[c], n0 + c) : 0ull;
                    topk_list_offer(rk, k, key, ok, lane);
                }
            }
        }
        __syncthreads();
    }
    if (k > 0) {
        for (int e = tid; e < TM * k; e += 256) {
            const int r = e / k, p = e % k;
            const int gr = row0 + r;
            if (gr < n_users) {
                const uint64_t key = keys[e];
                if (part_keys) {
                    part_keys[((int64_t)gr * gridDim.y + blockIdx.y) * k + p] = key;
                } else {
                    out_idx[(int64_t)gr * k + p] = key_id(key);
                    if (out_score) out_score[(int64_t)gr * k + p] = key_score(key);
                }
            }
        }
    }
}

// merges the per-split local top-k lists of a row (keys are globally comparable): one warp per row
__global__ void __launch_bounds__(256) merge_topk_kernel(const uint64_t *__restrict__ part_keys, int n_rows, int splits,
                                                         int k, int32_t *__restrict__ out_idx,
                                                         float *__restrict__ out_score) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t *rk = reinterpret_cast<uint64_t *>(smem_raw) + (size_t)wid * k;
    for (int row = blockIdx.x * 8 + wid; row < n_rows; row += gridDim.x * 8) {
        for (int p = lane; p < k; p += 32) rk[p] = make_key(-INFINITY, 0x7FFFFFFF);
        __syncwarp();
        const uint64_t *src = part_keys + (int64_t)row * splits * k;
        const int total = splits * k;
        for (int c0 = 0; c0 < total; c0 += 32) {
            const int c = c0 + lane;
            const bool ok = c < total;
            topk_list_offer(rk, k, ok ? src[c] : 0ull, ok, lane);
        }
        __syncwarp();
        for (int p = lane; p < k; p += 32) {
            out_idx[(int64_t)row * k + p] = key_id(rk[p]);
            if (out_score) out_score[(int64_t)row * k + p] = key_score(rk[p]);
        }
        __syncwarp();
    }
}

template <int TM>
static size_t exact_smem(int k) {
    using C = ExactCfg<TM>;
    size_t fl = (size_t)kPad * TM + (size_t)kPad * C::TN + (size_t)TM * (C::TN + 1);
    fl += fl & 1;
    return fl * 4 + (size_t)TM * 8 + (size_t)TM * (size_t)k * 8;
}

template <int TM>
static int launch_exact(const float *U, const float *V, int ld, int d, const int32_t *users, int n_users,
                        int num_items, const int64_t *mi, const int32_t *mx, int k, int32_t *oi, float *os,
                        float *dense, cudaStream_t s) {
    auto kern = score_topk_exact_kernel<TM>;
    const size_t smem = exact_smem<TM>(k);
    B200_REQUIRE(smem <= 227 * 1024, B200REC_EUNSUPPORTED, "score_topk: k=%d too large", k);
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (n_users + TM - 1) / TM;
    using C = ExactCfg<TM>;
    const int n_tiles = (num_items + C::TN - 1) / C::TN;
    const int sms = sm_count();
    // few row blocks (e.g. the TC path's overflow rows, small evaluations): split the catalogue so the
    // whole chip works on them, then merge the per-split top-k lists
    int splits = 1;
    if (k > 0 && !dense && grid < 2 * sms) {
        splits = (2 * sms + grid - 1) / grid;
        if (splits > n_tiles) splits = n_tiles;
        if (splits > 512) splits = 512;
    }
    if (splits <= 1) {
        kern<<<grid, 256, smem, s>>>(U, V, ld, d, users, n_users, num_items, mi, mx, k, oi, os, dense, num_items,
                                     nullptr);
        B200_LAUNCH_CHECK();
        return B200REC_OK;
    }
    const int tiles_per_split = (n_tiles + splits - 1) / splits;
    splits = (n_tiles + tiles_per_split - 1) / tiles_per_split;
    uint64_t *part = nullptr;
    {   // keep stream-ordered scratch cached across calls (the default pool trims to zero at every sync)
        static thread_local int pool_dev = -1;
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev != pool_dev) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                uint64_t keep = 1ull << 30;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            pool_dev = dev;
        }
    }
    B200_CUDA(cudaMallocAsync(&part, (size_t)n_users * splits * k * sizeof(uint64_t), s));
    kern<<<dim3(grid, splits), 256, smem, s>>>(U, V, ld, d, users, n_users, num_items, mi, mx, k, nullptr, nullptr,
                                               nullptr, tiles_per_split * C::TN, part);
    B200_LAUNCH_CHECK();
    const size_t msmem = (size_t)8 * k * 8;
    B200_CUDA(cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
    int mgrid = (n_users + 7) / 8;
    if (mgrid > sms * 8) mgrid = sms * 8;
    merge_topk_kernel<<<mgrid, 256, msmem, s>>>(part, n_users, splits, k, oi, os);
    B200_LAUNCH_CHECK();
    B200_CUDA(cudaFreeAsync(part, s));
    return B200REC_OK;
}

int score_topk_exact(const float *U, const float *V, int ld, int d, const int32_t *users, int n_users, int num_items,
                     const int64_t *mi, const int32_t *mx, int k, int32_t *oi, float *os, float *dense,
                     cudaStream_t s) {
    if (n_users <= 0) return B200REC_OK;
    if (k <= 256) return launch_exact<64>(U, V, ld, d, users, n_users, num_items, mi, mx, k, oi, os, dense, s);
    return launch_exact<16>(U, V, ld, d, users, n_users, num_items, mi, mx, k, oi, os, dense, s);
}

// device-resident c_top_k_array_index (func.h:22-31): one warp per row
__global__ void __launch_bounds__(256) topk_rows_kernel(const float *__restrict__ scores, int64_t row_stride,
                                                        int rows, int cols, int k, int32_t *__restrict__ out_idx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t *rk = reinterpret_cast<uint64_t *>(smem_raw) + (size_t)wid * k;
    for (int row = blockIdx.x * 8 + wid; row < rows; row += gridDim.x * 8) {
        for (int p = lane; p < k; p += 32) rk[p] = make_key(-INFINITY, 0x7FFFFFFF);
        __syncwarp();
        const float *sr = scores + (int64_t)row * row_stride;
        for (int c0 = 0; c0 < cols; c0 += 32) {
            const int c = c0 + lane;
            const bool ok = c < cols;
            const uint64_t key = ok ? make_key(sr[c], c) : 0ull;
            topk_list_offer(rk, k, key, ok, lane);
        }
        for (int p = lane; p < k; p += 32) out_idx[(int64_t)row * k + p] = key_id(rk[p]);
        __syncwarp();
    }
}

}  // namespace b200

using namespace b200;

extern "C" int b200rec_predict_dense(const float *U, const float *V, int ld, int d, const int32_t *users, int n_users,
                                     int num_items, const int64_t *mask_indptr, const int32_t *mask_indices,
                                     float *out, void *stream) {
    B200_REQUIRE(U && V && users && out, B200REC_EINVAL, "predict_dense: null argument");
    B200_REQUIRE(d >= 1 && ld >= d && num_items >= 1, B200REC_EINVAL, "predict_dense: bad sizes");
    B200_REQUIRE((mask_indptr == nullptr) == (mask_indices == nullptr), B200REC_EINVAL, "predict_dense: half a mask");
    return score_topk_exact(U, V, ld, d, users, n_users, num_items, mask_indptr, mask_indices, 0, nullptr, nullptr,
                            out, (cudaStream_t)stream);
}

extern "C" int b200rec_topk_rows(const float *scores, int64_t row_stride, int rows, int cols, int k,
                                 int32_t *out_idx, void *stream) {
    B200_REQUIRE(scores && out_idx, B200REC_EINVAL, "topk_rows: null argument");
    B200_REQUIRE(k >= 1 && k <= cols && k <= 1024 && row_stride >= cols, B200REC_EINVAL,
                 "topk_rows: need 1 <= k <= min(cols,1024) (k=%d cols=%d)", k, cols);
    if (rows <= 0) return B200REC_OK;
    const size_t smem = (size_t)8 * k * 8;
    B200_CUDA(cudaFuncSetAttribute(topk_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = (rows + 7) / 8;
    const int cap = sm_count() * 8;
    if (grid > cap) grid = cap;
    topk_rows_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(scores, row_stride, rows, cols, k, out_idx);
    B200_LAUNCH_CHECK();
    return B200REC_OK;
}
