// Library-level plumbing and the HOST-buffer drop-ins for the reference's native
// evaluation layer (evaluation/backend/cython/include/{func,holdout,loo}.h).
#include <atomic>
#include <string.h>
#include <vector>
#include "common.cuh"

namespace b200 {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
    static thread_local int cached_dev = -1, cached = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return cached;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
        cached_dev = dev;
    }
    return cached;
}

int score_topk_exact(const float *U, const float *V, int ld, int d, const int32_t *users, int n_users, int num_items,
                     const int64_t *mi, const int32_t *mx, int k, int32_t *oi, float *os, float *dense,
                     cudaStream_t s);
int score_topk_tc(const float *U, const float *V, int ld, int d, const int32_t *users, int n_users, int num_items,
                  const int64_t *mi, const int32_t *mx, int k, int32_t *oi, float *os, void *ws, int64_t ws_bytes,
                  cudaStream_t s);
int64_t score_topk_tc_workspace(int n_users, int num_items, int d, int k);

static int require_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n < 1) {
        set_error("no usable CUDA device (%s); libb200rec has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return B200REC_ECUDA;
    }
    return B200REC_OK;
}

struct DevBuf {  // RAII device scratch for the host-buffer drop-ins
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) {
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
        if (e != cudaSuccess) { set_error("cudaMalloc(%zu) -> %s", bytes, cudaGetErrorString(e)); p = nullptr; return B200REC_ENOMEM; }
        return B200REC_OK;
    }
    template <class T> T *as() { return reinterpret_cast<T *>(p); }
};

struct StreamPair {  // RAII: the two copy/compute streams of the chunked host drop-in (freed on every return path)
    cudaStream_t st[2] = {nullptr, nullptr};
    ~StreamPair() { for (int b = 0; b < 2; ++b) if (st[b]) cudaStreamDestroy(st[b]); }
    int create() {
        for (int b = 0; b < 2; ++b) {
            cudaError_t e = cudaStreamCreate(&st[b]);
            if (e != cudaSuccess) { set_error("cudaStreamCreate -> %s", cudaGetErrorString(e)); st[b] = nullptr; return B200REC_ECUDA; }
        }
        return B200REC_OK;
    }
};

}  // namespace b200

using namespace b200;

extern "C" const char *b200rec_last_error(void) { return g_err; }
extern "C" int b200rec_version(void) { return 100; }
extern "C" int64_t b200rec_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int64_t b200rec_score_topk_workspace(int n_users, int num_items, int d, int k, int algo) {
    if (algo == B200REC_SCORE_TC) return score_topk_tc_workspace(n_users, num_items, d, k);
    return 0;
}

extern "C" int b200rec_score_topk(const float *U, const float *V, int ld, int d, const int32_t *users, int n_users,
                                  int num_items, const int64_t *mask_indptr, const int32_t *mask_indices, int k,
                                  int32_t *out_idx, float *out_score, void *workspace, int64_t workspace_bytes,
                                  int algo, void *stream) {
    B200_REQUIRE(U && V && users && out_idx, B200REC_EINVAL, "score_topk: null argument");
    B200_REQUIRE(d >= 1 && ld >= d && num_items >= 1 && n_users >= 0, B200REC_EINVAL, "score_topk: bad sizes");
    B200_REQUIRE(k >= 1 && k <= num_items && k <= 1024, B200REC_EINVAL,
                 "score_topk: need 1 <= k <= min(num_items, 1024) (k=%d, num_items=%d)", k, num_items);
    B200_REQUIRE((mask_indptr == nullptr) == (mask_indices == nullptr), B200REC_EINVAL, "score_topk: half a mask");
    if (algo == B200REC_SCORE_TC)
        return score_topk_tc(U, V, ld, d, users, n_users, num_items, mask_indptr, mask_indices, k, out_idx, out_score,
                             workspace, workspace_bytes, (cudaStream_t)stream);
    B200_REQUIRE(algo == B200REC_SCORE_EXACT, B200REC_EINVAL, "score_topk: unknown algo %d", algo);
    return score_topk_exact(U, V, ld, d, users, n_users, num_items, mask_indptr, mask_indices, k, out_idx, out_score,
                            nullptr, (cudaStream_t)stream);
}

// Reserve part of L2 for a reused table (the item table of the BPR step: 51 MB at cfg2, read twice and
// vector-reduced twice per triple) while 1 GB of single-use user rows stream past it: sets the device's persisting-L2
// carve-out and an access-policy window on `stream` (kernels launched on it afterwards).  bytes == 0 clears it.
extern "C" int b200rec_l2_persist(const void *base, int64_t bytes, float hit_ratio, void *stream) {
    int rc = require_device();
    if (rc) return rc;
    int dev = 0, max_persist = 0, max_window = 0;
    B200_CUDA(cudaGetDevice(&dev));
    B200_CUDA(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev));
    B200_CUDA(cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev));
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof(attr));
    if (bytes <= 0 || base == nullptr) {
        attr.accessPolicyWindow.num_bytes = 0;
        B200_CUDA(cudaStreamSetAttribute((cudaStream_t)stream, cudaStreamAttributeAccessPolicyWindow, &attr));
        B200_CUDA(cudaCtxResetPersistingL2Cache());
        return B200REC_OK;
    }
    B200_REQUIRE(max_persist > 0, B200REC_EUNSUPPORTED, "l2_persist: device has no persisting L2");
    size_t carve = (size_t)bytes < (size_t)max_persist ? (size_t)bytes : (size_t)max_persist;
    B200_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve));
    size_t win = (size_t)bytes < (size_t)max_window ? (size_t)bytes : (size_t)max_window;
    attr.accessPolicyWindow.base_ptr = const_cast<void *>(base);
    attr.accessPolicyWindow.num_bytes = win;
    attr.accessPolicyWindow.hitRatio = hit_ratio * (win > carve ? (float)carve / (float)win : 1.f);
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    B200_CUDA(cudaStreamSetAttribute((cudaStream_t)stream, cudaStreamAttributeAccessPolicyWindow, &attr));
    return B200REC_OK;
}

// ---- func.h:22-31 --------------------------------------------------------
extern "C" int b200rec_top_k_array_index(const float *scores_pt, int columns_num, int rows_num, int max_k,
                                         int *rankings_pt) {
    B200_REQUIRE(scores_pt && rankings_pt, B200REC_EINVAL, "top_k_array_index: null argument");
    B200_REQUIRE(columns_num >= 1 && rows_num >= 0 && max_k >= 1 && max_k <= columns_num && max_k <= 1024,
                 B200REC_EINVAL, "top_k_array_index: need 1 <= max_k <= min(columns_num,1024) (max_k=%d cols=%d)", max_k,
                 columns_num);
    if (rows_num == 0) return B200REC_OK;
    int rc = require_device();
    if (rc) return rc;
    // stream the score matrix through the device in row chunks of <= 256 MiB, double-buffered on two streams
    const size_t row_bytes = (size_t)columns_num * sizeof(float);
    int chunk_rows = (int)((256ull << 20) / row_bytes);
    if (chunk_rows < 1) chunk_rows = 1;
    if (chunk_rows > rows_num) chunk_rows = rows_num;
    StreamPair sp;
    if ((rc = sp.create())) return rc;
    cudaStream_t *st = sp.st;
    DevBuf sc[2], ix[2];
    for (int b = 0; b < 2; ++b) {
        if ((rc = sc[b].alloc((size_t)chunk_rows * row_bytes))) return rc;
        if ((rc = ix[b].alloc((size_t)chunk_rows * max_k * sizeof(int)))) return rc;
    }
    int b = 0;
    for (int r0 = 0; r0 < rows_num; r0 += chunk_rows, b ^= 1) {
        const int nr = (rows_num - r0) < chunk_rows ? (rows_num - r0) : chunk_rows;
        B200_CUDA(cudaMemcpyAsync(sc[b].p, scores_pt + (size_t)r0 * columns_num, (size_t)nr * row_bytes,
                                  cudaMemcpyHostToDevice, st[b]));
        rc = b200rec_topk_rows(sc[b].as<float>(), columns_num, nr, columns_num, max_k, ix[b].as<int32_t>(), st[b]);
        if (rc) return rc;
        B200_CUDA(cudaMemcpyAsync(rankings_pt + (size_t)r0 * max_k, ix[b].p, (size_t)nr * max_k * sizeof(int),
                                  cudaMemcpyDeviceToHost, st[b]));
    }
    B200_CUDA(cudaStreamSynchronize(st[0]));
    B200_CUDA(cudaStreamSynchronize(st[1]));
    return B200REC_OK;
}

// flatten the reference's int** ground-truth table (holdout_func.pyx:22-30) into CSR on the device
static int upload_truth(int users_num, int **ground_truths, const int *lens, DevBuf &dptr, DevBuf &didx) {
    std::vector<int64_t> ptr((size_t)users_num + 1, 0);
    for (int u = 0; u < users_num; ++u) ptr[u + 1] = ptr[u] + (lens ? lens[u] : 1);
    std::vector<int32_t> idx((size_t)ptr[users_num]);
    for (int u = 0; u < users_num; ++u)
        memcpy(idx.data() + ptr[u], ground_truths[u], sizeof(int32_t) * (size_t)(ptr[u + 1] - ptr[u]));
    int rc;
    if ((rc = dptr.alloc(ptr.size() * sizeof(int64_t)))) return rc;
    if ((rc = didx.alloc(idx.size() * sizeof(int32_t)))) return rc;
    B200_CUDA(cudaMemcpy(dptr.p, ptr.data(), ptr.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
    B200_CUDA(cudaMemcpy(didx.p, idx.data(), idx.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    return B200REC_OK;
}

// ---- holdout.h:20-103 ----------------------------------------------------
extern "C" int b200rec_evaluate_holdout(int users_num, const int *rankings, int max_k, const int *Ks, int K_len,
                                        int **ground_truths, const int *ground_truths_num, float *results) {
    B200_REQUIRE(rankings && Ks && ground_truths && ground_truths_num && results, B200REC_EINVAL,
                 "evaluate_holdout: null argument");
    if (users_num <= 0) return B200REC_OK;
    int rc = require_device();
    if (rc) return rc;
    DevBuf dptr, didx, dtop, dout;
    if ((rc = upload_truth(users_num, ground_truths, ground_truths_num, dptr, didx))) return rc;
    if ((rc = dtop.alloc((size_t)users_num * max_k * sizeof(int)))) return rc;
    if ((rc = dout.alloc((size_t)users_num * 3 * K_len * sizeof(float)))) return rc;
    B200_CUDA(cudaMemcpy(dtop.p, rankings, (size_t)users_num * max_k * sizeof(int), cudaMemcpyHostToDevice));
    rc = b200rec_holdout_metrics(dtop.as<int32_t>(), users_num, max_k, nullptr, dptr.as<int64_t>(), didx.as<int32_t>(),
                                 Ks, K_len, dout.as<float>(), nullptr);
    if (rc) return rc;
    B200_CUDA(cudaMemcpy(results, dout.p, (size_t)users_num * 3 * K_len * sizeof(float), cudaMemcpyDeviceToHost));
    return B200REC_OK;
}

// ---- loo.h:20-85 ---------------------------------------------------------
extern "C" int b200rec_evaluate_loo(int users_num, const int *rankings, int max_k, const int *Ks, int K_len,
                                    int **ground_truths, float *results) {
    B200_REQUIRE(rankings && Ks && ground_truths && results, B200REC_EINVAL, "evaluate_loo: null argument");
    if (users_num <= 0) return B200REC_OK;
    int rc = require_device();
    if (rc) return rc;
    DevBuf dptr, didx, dtop, dout;
    if ((rc = upload_truth(users_num, ground_truths, nullptr, dptr, didx))) return rc;
    if ((rc = dtop.alloc((size_t)users_num * max_k * sizeof(int)))) return rc;
    if ((rc = dout.alloc((size_t)users_num * 2 * K_len * sizeof(float)))) return rc;
    B200_CUDA(cudaMemcpy(dtop.p, rankings, (size_t)users_num * max_k * sizeof(int), cudaMemcpyHostToDevice));
    rc = b200rec_loo_metrics(dtop.as<int32_t>(), users_num, max_k, nullptr, dptr.as<int64_t>(), didx.as<int32_t>(), Ks,
                             K_len, dout.as<float>(), nullptr);
    if (rc) return rc;
    B200_CUDA(cudaMemcpy(results, dout.p, (size_t)users_num * 2 * K_len * sizeof(float), cudaMemcpyDeviceToHost));
    return B200REC_OK;
}
