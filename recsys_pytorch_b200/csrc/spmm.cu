// CSR SpMM for LightGCN propagation (sm_100a):  Y = A X,  acc += scale * Y.
//
// Replaces models/LightGCN.py:196 `torch.sparse.mm(g_droped, all_emb)` (cuSPARSE
// COO SpMM) and the stack+mean of :198-200 (fused here as a running accumulation).
// HBM-bound gather: one sub-group of G lanes per output row (G*16 B >= row bytes
// when d <= 128); (col,val) pairs are loaded coalesced by the sub-group and
// broadcast with shuffles; 4 neighbour rows in flight per lane.
#include "common.cuh"

namespace b200 {

// a[] += sum over nnz [start, end) of values * X[indices] for the sub-group's row piece
template <int G, int CPL>
__device__ __forceinline__ void accumulate_range(const int32_t *__restrict__ indices, const float *__restrict__ values,
                                                 int64_t start, int64_t end, const float *__restrict__ X, int ldx, int d4,
                                                 int sl, int sg, unsigned gmask, float4 (&a)[CPL]) {
    for (int64_t base = start; base < end; base += G) {
        const int64_t mp = base + sl;
        int c = 0;
        float v = 0.f;
        if (mp < end) { c = indices[mp]; v = values[mp]; }
        const int cnt = (int)((end - base) < G ? (end - base) : G);
        // two-level summation: the <= G products of this block go into b[], then b[] into a[] - keeps the
        // fp32 error of 100k-neighbour rows (popular items) at (deg/G) eps instead of deg eps; neighbour
        // rows are fetched four at a time so four gathers are in flight per lane
        float4 b[CPL];
#pragma unroll
        for (int k = 0; k < CPL; ++k) b[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int t = 0; t < cnt; t += 4) {
            int cc[4];
            float vv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int src = sg * G + ((t + i < G) ? (t + i) : (G - 1));
                cc[i] = __shfl_sync(gmask, c, src);
                vv[i] = __shfl_sync(gmask, v, src);
                if (t + i >= cnt) { vv[i] = 0.f; cc[i] = cc[0]; }
            }
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                const int q = sl + k * G;
                if (q < d4) {
                    float4 x[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) x[i] = ld4(X + (int64_t)cc[i] * ldx + q * 4);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        b[k].x = fmaf(vv[i], x[i].x, b[k].x); b[k].y = fmaf(vv[i], x[i].y, b[k].y);
                        b[k].z = fmaf(vv[i], x[i].z, b[k].z); b[k].w = fmaf(vv[i], x[i].w, b[k].w);
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < CPL; ++k) { a[k].x += b[k].x; a[k].y += b[k].y; a[k].z += b[k].z; a[k].w += b[k].w; }
    }
}

// row result -> Y and/or the running layer mean
template <int G, int CPL>
__device__ __forceinline__ void store_row(int64_t row, const float4 (&a)[CPL], int sl, int d4, const float *__restrict__ X,
                                          int ldx, float *__restrict__ Y, int ldy, float *__restrict__ acc, int ldacc,
                                          float acc_scale, int acc_init) {
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
        const int q = sl + k * G;
        if (q < d4) {
            if (Y) st4(Y + row * ldy + q * 4, a[k]);
            if (acc) {
                float *pa = acc + row * ldacc + q * 4;
                float4 o = ld4(pa);
                if (acc_init) {  // first layer: running mean starts as scale * E_0 (models/LightGCN.py:198-200)
                    const float4 x0 = ld4(X + row * ldx + q * 4);
                    o = make_float4(acc_scale * x0.x, acc_scale * x0.y, acc_scale * x0.z, acc_scale * x0.w);
                }
                o.x = fmaf(acc_scale, a[k].x, o.x); o.y = fmaf(acc_scale, a[k].y, o.y);
                o.z = fmaf(acc_scale, a[k].z, o.z); o.w = fmaf(acc_scale, a[k].w, o.w);
                st4(pa, o);
            }
        }
    }
}

// rows with more than `skip_above` neighbours are left to the split path (skip_above < 0: no row is skipped)
template <int G, int CPL>
__global__ void __launch_bounds__(256) spmm_csr_kernel(const int64_t *__restrict__ indptr,
                                                       const int32_t *__restrict__ indices,
                                                       const float *__restrict__ values, int n_rows,
                                                       const float *__restrict__ X, int ldx, int d4,
                                                       float *__restrict__ Y, int ldy, float *__restrict__ acc,
                                                       int ldacc, float acc_scale, int acc_init, int64_t skip_above) {
    constexpr int RPW = 32 / G;
    const int lane = threadIdx.x & 31;
    const int sl = lane % G, sg = lane / G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (sg * G));
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t rb = warp_global * RPW; rb < n_rows; rb += n_warps * RPW) {
        const int64_t row = rb + sg;
        if (row >= n_rows) continue;  // whole sub-group leaves together
        const int64_t start = indptr[row], end = indptr[row + 1];
        if (skip_above >= 0 && end - start > skip_above) continue;
        float4 a[CPL];
#pragma unroll
        for (int k = 0; k < CPL; ++k) a[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        accumulate_range<G, CPL>(indices, values, start, end, X, ldx, d4, sl, sg, gmask, a);
        store_row<G, CPL>(row, a, sl, d4, X, ldx, Y, ldy, acc, ldacc, acc_scale, acc_init);
    }
}

// Split path for long rows (popular items: up to ~1M neighbours at cfg4 - one sub-group walking such a row alone
// took 178 ms per layer): every long row is cut into segments of <= seg_len nnz, one sub-group per segment writes
// a partial row, and one sub-group per long row adds its partials IN ORDER (deterministic) and finishes the row.
template <int G, int CPL>
__global__ void __launch_bounds__(256) spmm_segment_kernel(const int32_t *__restrict__ indices,
                                                           const float *__restrict__ values,
                                                           const int64_t *__restrict__ seg_begin,
                                                           const int64_t *__restrict__ seg_end, int n_seg,
                                                           const float *__restrict__ X, int ldx, int d4,
                                                           float *__restrict__ partial, int ldp) {
    constexpr int RPW = 32 / G;
    const int lane = threadIdx.x & 31;
    const int sl = lane % G, sg = lane / G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (sg * G));
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t sb = warp_global * RPW; sb < n_seg; sb += n_warps * RPW) {
        const int64_t seg = sb + sg;
        if (seg >= n_seg) continue;
        float4 a[CPL];
#pragma unroll
        for (int k = 0; k < CPL; ++k) a[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        accumulate_range<G, CPL>(indices, values, seg_begin[seg], seg_end[seg], X, ldx, d4, sl, sg, gmask, a);
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            const int q = sl + k * G;
            if (q < d4) st4(partial + seg * ldp + q * 4, a[k]);
        }
    }
}

template <int G, int CPL>
__global__ void __launch_bounds__(256) spmm_long_rows_kernel(const int32_t *__restrict__ long_rows,
                                                             const int32_t *__restrict__ long_seg_ptr, int n_long,
                                                             const float *__restrict__ partial, int ldp,
                                                             const float *__restrict__ X, int ldx, int d4,
                                                             float *__restrict__ Y, int ldy, float *__restrict__ acc,
                                                             int ldacc, float acc_scale, int acc_init) {
    constexpr int RPW = 32 / G;
    const int lane = threadIdx.x & 31;
    const int sl = lane % G, sg = lane / G;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t rb = warp_global * RPW; rb < n_long; rb += n_warps * RPW) {
        const int64_t r = rb + sg;
        if (r >= n_long) continue;
        float4 a[CPL];
#pragma unroll
        for (int k = 0; k < CPL; ++k) a[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int sgm = long_seg_ptr[r]; sgm < long_seg_ptr[r + 1]; ++sgm) {
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                const int q = sl + k * G;
                if (q < d4) {
                    const float4 x = ld4(partial + (int64_t)sgm * ldp + q * 4);
                    a[k].x += x.x; a[k].y += x.y; a[k].z += x.z; a[k].w += x.w;
                }
            }
        }
        store_row<G, CPL>((int64_t)long_rows[r], a, sl, d4, X, ldx, Y, ldy, acc, ldacc, acc_scale, acc_init);
    }
}

struct SpmmSplit {   // optional plan for long rows (all device pointers); n_seg == 0: plain path
    int64_t seg_len;
    const int64_t *seg_begin, *seg_end;
    int n_seg;
    const int32_t *long_rows, *long_seg_ptr;
    int n_long;
    float *partial;
};

template <int G, int CPL>
static int launch_spmm(const int64_t *indptr, const int32_t *indices, const float *values, int n_rows, const float *X,
                       int ldx, int d4, float *Y, int ldy, float *acc, int ldacc, float sc, int acc_init,
                       const SpmmSplit &sp, cudaStream_t s) {
    constexpr int RPW = 32 / G;
    int64_t blocks = ((int64_t)n_rows + 8 * RPW - 1) / (8 * RPW);
    const int64_t cap = (int64_t)sm_count() * 8;
    spmm_csr_kernel<G, CPL><<<(int)(blocks < cap ? blocks : cap), 256, 0, s>>>(indptr, indices, values, n_rows, X, ldx,
                                                                               d4, Y, ldy, acc, ldacc, sc, acc_init,
                                                                               sp.n_seg > 0 ? sp.seg_len : (int64_t)-1);
    B200_LAUNCH_CHECK();
    if (sp.n_seg > 0) {
        const int ldp = d4 * 4;
        blocks = ((int64_t)sp.n_seg + 8 * RPW - 1) / (8 * RPW);
        spmm_segment_kernel<G, CPL><<<(int)(blocks < cap ? blocks : cap), 256, 0, s>>>(indices, values, sp.seg_begin,
                                                                                      sp.seg_end, sp.n_seg, X, ldx, d4,
                                                                                      sp.partial, ldp);
        B200_LAUNCH_CHECK();
        blocks = ((int64_t)sp.n_long + 8 * RPW - 1) / (8 * RPW);
        spmm_long_rows_kernel<G, CPL><<<(int)(blocks < cap ? blocks : cap), 256, 0, s>>>(
            sp.long_rows, sp.long_seg_ptr, sp.n_long, sp.partial, ldp, X, ldx, d4, Y, ldy, acc, ldacc, sc, acc_init);
        B200_LAUNCH_CHECK();
    }
    return B200REC_OK;
}

}  // namespace b200

using namespace b200;

static int spmm_dispatch(const int64_t *indptr, const int32_t *indices, const float *values, int n_rows, const float *X,
                         int ldx, int d, float *Y, int ldy, float *acc, int ldacc, float acc_scale, int acc_init,
                         const SpmmSplit &sp, void *stream) {
    B200_REQUIRE(indptr && indices && values && X && (Y || acc), B200REC_EINVAL, "spmm_csr: null argument");
    B200_REQUIRE(d >= 1 && ldx >= d && ldx % 4 == 0 && ldx <= 512, B200REC_EINVAL, "spmm_csr: bad d/ldx");
    B200_REQUIRE((!Y || (ldy >= d && ldy % 4 == 0)) && (!acc || (ldacc >= d && ldacc % 4 == 0)), B200REC_EINVAL,
                 "spmm_csr: bad ldy/ldacc");
    B200_REQUIRE(X != Y, B200REC_EINVAL, "spmm_csr: in-place propagation is not supported");
    if (n_rows <= 0) return B200REC_OK;
    const int d4 = (d + 3) / 4;
    int G = 1;
    while (G < d4 && G < 32) G <<= 1;
    const int CPL = (d4 + G - 1) / G;
    cudaStream_t s = (cudaStream_t)stream;
#define B200_SPMM(GG, CC) return launch_spmm<GG, CC>(indptr, indices, values, n_rows, X, ldx, d4, Y, ldy, acc, ldacc, acc_scale, acc_init, sp, s)
    switch (G) {
        case 1: B200_SPMM(1, 1);
        case 2: B200_SPMM(2, 1);
        case 4: B200_SPMM(4, 1);
        case 8: B200_SPMM(8, 1);
        case 16: B200_SPMM(16, 1);
        default:
            switch (CPL) {
                case 1: B200_SPMM(32, 1);
                case 2: B200_SPMM(32, 2);
                case 3: B200_SPMM(32, 3);
                default: B200_SPMM(32, 4);
            }
    }
#undef B200_SPMM
}

extern "C" int b200rec_spmm_csr(const int64_t *indptr, const int32_t *indices, const float *values, int n_rows,
                                const float *X, int ldx, int d, float *Y, int ldy, float *acc, int ldacc,
                                float acc_scale, int acc_init, void *stream) {
    SpmmSplit sp = {};
    return spmm_dispatch(indptr, indices, values, n_rows, X, ldx, d, Y, ldy, acc, ldacc, acc_scale, acc_init, sp, stream);
}

extern "C" int b200rec_spmm_csr_split(const int64_t *indptr, const int32_t *indices, const float *values, int n_rows,
                                      const float *X, int ldx, int d, float *Y, int ldy, float *acc, int ldacc,
                                      float acc_scale, int acc_init, int64_t seg_len, const int64_t *seg_begin,
                                      const int64_t *seg_end, int n_seg, const int32_t *long_rows,
                                      const int32_t *long_seg_ptr, int n_long, float *partial, void *stream) {
    SpmmSplit sp = {};
    if (n_seg > 0) {
        B200_REQUIRE(seg_len >= 1 && seg_begin && seg_end && long_rows && long_seg_ptr && n_long >= 1 && partial,
                     B200REC_EINVAL, "spmm_csr_split: incomplete plan");
        sp.seg_len = seg_len; sp.seg_begin = seg_begin; sp.seg_end = seg_end; sp.n_seg = n_seg;
        sp.long_rows = long_rows; sp.long_seg_ptr = long_seg_ptr; sp.n_long = n_long; sp.partial = partial;
    }
    return spmm_dispatch(indptr, indices, values, n_rows, X, ldx, d, Y, ldy, acc, ldacc, acc_scale, acc_init, sp, stream);
}
