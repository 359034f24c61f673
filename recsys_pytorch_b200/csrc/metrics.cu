// Holdout / leave-one-out ranking metrics on the device (sm_100a).
//
// Replaces evaluation/backend/cython/include/holdout.h:20-103 and loo.h:20-85
// together with the per-user Python loop of evaluation/backend/cython/
// holdout.py:23-27.  Arithmetic is kept bit-identical to the
// C++: `float` accumulators, each discount 1.0/log2(i+2) a double computed on the
// HOST with the same libm call and added in double, then rounded to float.
#include <math.h>
#include <vector>
#include "common.cuh"

namespace b200 {

constexpr int kMaxK = 1024;
__constant__ double c_inv_log2[kMaxK + 2];  // [i] = 1.0 / log2(i + 2); per DEVICE (constant memory is), uploaded once each

struct KsArg {      // the cut-offs travel as a kernel argument: no shared constant buffer that a later call on another
    int n;          // stream could overwrite while an earlier kernel still reads it
    int v[64];
};

// discount table: computed once on the host with the same libm call the reference uses (holdout.h:47, loo.h:51), kept in
// static storage so the asynchronous upload needs no synchronisation, uploaded once per device
static int upload_tables(int max_k, cudaStream_t s) {
    static double table[kMaxK + 2];
    static bool table_ready = false;
    static bool uploaded[64] = {false};
    (void)max_k;
    if (!table_ready) {
        for (int i = 0; i < kMaxK + 2; ++i) table[i] = 1.0 / log2((double)(i + 2));
        table_ready = true;
    }
    int dev = 0;
    B200_CUDA(cudaGetDevice(&dev));
    B200_REQUIRE(dev >= 0 && dev < 64, B200REC_EINVAL, "metrics: device index %d", dev);
    if (!uploaded[dev]) {
        B200_CUDA(cudaMemcpyToSymbolAsync(c_inv_log2, table, sizeof(double) * (kMaxK + 2), 0, cudaMemcpyHostToDevice, s));
        // later calls may use OTHER streams: make the table visible device-wide before anyone reads it (once per device)
        B200_CUDA(cudaStreamSynchronize(s));
        uploaded[dev] = true;
    }
    return B200REC_OK;
}

static KsArg make_ks(const int *Ks, int K_len) {
    KsArg k;
    k.n = K_len;
    for (int j = 0; j < 64; ++j) k.v[j] = j < K_len ? Ks[j] : 0;
    return k;
}

__device__ __forceinline__ bool in_truth(const int32_t *truth, int n, int v) {
    for (int i = 0; i < n; ++i)
        if (truth[i] == v) return true;  // std::set membership (holdout.h:37,43); rows are short
    return false;
}

// one thread per user (holdout.h:29-66)
__global__ void __launch_bounds__(128) holdout_kernel(const int32_t *__restrict__ topk, int n, int max_k,
                                                      const int32_t *__restrict__ row_ids,
                                                      const int64_t *__restrict__ tptr,
                                                      const int32_t *__restrict__ tidx, const KsArg ks,
                                                      float *__restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int K_len = ks.n;
    const int64_t row = row_ids ? row_ids[r] : r;
    const int32_t *truth = tidx + tptr[row];
    const int truth_len = (int)(tptr[row + 1] - tptr[row]);
    const int32_t *cur = topk + (int64_t)r * max_k;
    float *res = out + (int64_t)r * 3 * K_len;
    float hits = 0.f, iDCG = 0.f, DCG = 0.f;
    for (int i = 0; i < max_k; ++i) {
        if (in_truth(truth, truth_len, cur[i])) {
            hits += 1.f;
            DCG = (float)((double)DCG + c_inv_log2[i]);
        }
        if (i < truth_len) iDCG = (float)((double)iDCG + c_inv_log2[i]);
        for (int j = 0; j < K_len; ++j)
            if (ks.v[j] == i + 1) {
                res[j] = hits / (float)ks.v[j];
                res[K_len + j] = hits / (float)truth_len;
                res[2 * K_len + j] = DCG / iDCG;
            }
    }
}

// loo.h:29-62
__global__ void __launch_bounds__(128) loo_kernel(const int32_t *__restrict__ topk, int n, int max_k,
                                                  const int32_t *__restrict__ row_ids,
                                                  const int64_t *__restrict__ tptr,
                                                  const int32_t *__restrict__ tidx, const KsArg ks,
                                                  float *__restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int K_len = ks.n;
    const int64_t row = row_ids ? row_ids[r] : r;
    const int32_t truth = tidx[tptr[row]];
    const int32_t *cur = topk + (int64_t)r * max_k;
    float *res = out + (int64_t)r * 2 * K_len;
    int hit_at_k = max_k + 1;
    for (int i = 0; i < max_k; ++i)
        if (cur[i] == truth) { hit_at_k = i + 1; break; }
    for (int j = 0; j < K_len; ++j) {
        if (ks.v[j] >= hit_at_k) {
            res[j] = 1.0f;
            res[K_len + j] = (float)c_inv_log2[hit_at_k - 1];  // 1 / log2(hit_at_k + 1)
        } else {
            res[j] = 0.f;
            res[K_len + j] = 0.f;
        }
    }
}

// utils/stats.py:29-32 column means; fp64 accumulation (numpy's pairwise fp32 sum agrees to ~1e-7)
__global__ void __launch_bounds__(256) colsum_kernel(const float *__restrict__ mat, int64_t n, int cols,
                                                     double *__restrict__ sums) {
    __shared__ double sh[256];
    for (int c = 0; c < cols; ++c) {
        double acc = 0.0;
        for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
            acc += (double)mat[r * cols + c];
        sh[threadIdx.x] = acc;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) atomicAdd(sums + c, sh[0]);
        __syncthreads();
    }
}

static int check_metric_args(const int32_t *topk, int n, int max_k, const int64_t *tp, const int32_t *ti,
                             const int *Ks, int K_len, float *out) {
    B200_REQUIRE(topk && tp && ti && Ks && out, B200REC_EINVAL, "metrics: null argument");
    B200_REQUIRE(n >= 0 && max_k >= 1 && max_k <= kMaxK && K_len >= 1 && K_len <= 64, B200REC_EINVAL,
                 "metrics: need 1 <= max_k <= %d, 1 <= K_len <= 64", kMaxK);
    return B200REC_OK;
}

}  // namespace b200

using namespace b200;

extern "C" int b200rec_holdout_metrics(const int32_t *topk, int n, int max_k, const int32_t *row_ids,
                                       const int64_t *truth_indptr, const int32_t *truth_indices, const int *Ks,
                                       int K_len, float *out, void *stream) {
    int rc = check_metric_args(topk, n, max_k, truth_indptr, truth_indices, Ks, K_len, out);
    if (rc) return rc;
    if (n == 0) return B200REC_OK;
    cudaStream_t s = (cudaStream_t)stream;
    rc = upload_tables(max_k, s);
    if (rc) return rc;
    B200_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)n * 3 * K_len, s));  // np.zeros (holdout_func.pyx:37)
    holdout_kernel<<<(n + 127) / 128, 128, 0, s>>>(topk, n, max_k, row_ids, truth_indptr, truth_indices,
                                                   make_ks(Ks, K_len), out);
    B200_LAUNCH_CHECK();
    return B200REC_OK;
}

extern "C" int b200rec_loo_metrics(const int32_t *topk, int n, int max_k, const int32_t *row_ids,
                                   const int64_t *truth_indptr, const int32_t *truth_indices, const int *Ks,
                                   int K_len, float *out, void *stream) {
    int rc = check_metric_args(topk, n, max_k, truth_indptr, truth_indices, Ks, K_len, out);
    if (rc) return rc;
    if (n == 0) return B200REC_OK;
    cudaStream_t s = (cudaStream_t)stream;
    rc = upload_tables(max_k, s);
    if (rc) return rc;
    loo_kernel<<<(n + 127) / 128, 128, 0, s>>>(topk, n, max_k, row_ids, truth_indptr, truth_indices, make_ks(Ks, K_len),
                                               out);
    B200_LAUNCH_CHECK();
    return B200REC_OK;
}

extern "C" int b200rec_column_means(const float *mat, int64_t n, int cols, double *out_host, void *stream) {
    B200_REQUIRE(mat && out_host && cols >= 1 && cols <= 4096 && n >= 1, B200REC_EINVAL, "column_means: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    // per-device scratch kept for the life of the process: a cudaMallocAsync/cudaFreeAsync pair here costs ~1 ms to
    // map and ~2 ms to trim at the caller's next device synchronisation (measured), for 32 KB
    // (the call ends with a synchronisation of `s`, and one host thread drives a device: the buffer is free again when
    // the next call on any stream of this device starts)
    static double *scratch[64] = {nullptr};
    int dev = 0;
    B200_CUDA(cudaGetDevice(&dev));
    B200_REQUIRE(dev >= 0 && dev < 64, B200REC_EINVAL, "column_means: device index %d", dev);
    if (!scratch[dev]) B200_CUDA(cudaMalloc(&scratch[dev], sizeof(double) * 4096));
    double *d_sums = scratch[dev];
    B200_CUDA(cudaMemsetAsync(d_sums, 0, sizeof(double) * cols, s));
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 4;
    colsum_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, s>>>(mat, n, cols, d_sums);
    B200_LAUNCH_CHECK();
    B200_CUDA(cudaMemcpyAsync(out_host, d_sums, sizeof(double) * cols, cudaMemcpyDeviceToHost, s));
    B200_CUDA(cudaStreamSynchronize(s));
    for (int c = 0; c < cols; ++c) out_host[c] /= (double)n;
    return B200REC_OK;
}
