// NGCF propagation layer for sm_100a.
//
// Replaces, per layer, models/NGCF.py:198-212 of the reference
//     sum_emb = side @ W_gc + b_gc;  bi_emb = (ego * side) @ W_bi + b_bi
//     ego'    = leaky_relu(sum_emb + bi_emb, 0.2);  ego' = dropout(ego', mess_dropout)
//     norm    = F.normalize(ego', p=2, dim=1);  embs += [norm]            (and the running mean of :216-218)
// (side = A_hat @ ego comes from the CSR SpMM of spmm.cu) and the autograd backward of exactly that, as three kernels:
//   ngcf_fwd_kernel    one warp per node row: the two [d,d] transforms from shared memory, activation, dropout mask from
//                      the counter RNG, L2 normalisation, running layer mean - one pass over the row
//   ngcf_bwd_row_kernel   per row: normalise / dropout / leaky-relu backward -> gz, then gz W^T products ->
//                      d(side) and the direct part of d(ego)           (d(ego) += A_hat d(side) is one more SpMM)
//   ngcf_wgrad_kernel  dW_gc = side^T gz, dW_bi = (ego*side)^T gz, db = sum gz: register-tiled over row chunks,
//                      one atomicAdd per output and CTA
// The transforms are [N,d]x[d,d] with d <= 64: 8k FMA per row against ~1 KB of row traffic - HBM-bound work for the
// CUDA cores, not tensor-core shaped (N = 1.1M rows at cfg4 is 18 GFLOP per layer, 30 us of FMA time).
#include "common.cuh"

namespace b200 {

constexpr int kNgcfMaxD = 64;

__device__ __forceinline__ bool ngcf_keep(uint64_t seed, uint64_t step, int layer, int64_t row, int col, float p) {
    if (p <= 0.f) return true;
    const uint32_t r = rng_u32(seed ^ (0x9E37ull * (uint64_t)(layer + 1)), step, (uint64_t)row * 64u + (uint64_t)col, 7);
    return (float)(r >> 8) * (1.0f / 16777216.0f) >= p;      // torch: keep with probability 1-p
}

struct NgcfFwd {
    const float *ego, *side;      // [N, ld]
    const float *Wg, *bg, *Wb, *bb;  // [d,d] row-major (in x out), [d]
    float *ego_next;              // [N, ld]  post-activation, post-dropout (input of the next layer)
    float *nrm;                   // [N]      max(|ego_next|, 1e-12)
    float *acc;                   // [N, ld]  running mean of the layer outputs
    float acc_scale;              // 1 / (L + 1)
    int N, ld, d, layer;
    float p_drop; uint64_t seed, step;
};

__global__ void __launch_bounds__(256) ngcf_fwd_kernel(const NgcfFwd a) {
    extern __shared__ float sm[];
    const int d = a.d;
    float *Wg = sm, *Wb = sm + d * d, *bs = Wb + d * d, *rowbuf = bs + d;   // rowbuf: [8 warps][2][d]
    for (int e = threadIdx.x; e < d * d; e += blockDim.x) { Wg[e] = a.Wg[e]; Wb[e] = a.Wb[e]; }
    for (int e = threadIdx.x; e < d; e += blockDim.x) bs[e] = a.bg[e] + a.bb[e];
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float *sb = rowbuf + wid * 2 * d, *tb = sb + d;
    const float keep_scale = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
    for (int64_t row = (int64_t)blockIdx.x * 8 + wid; row < a.N; row += (int64_t)gridDim.x * 8) {
        const float *pe = a.ego + row * a.ld, *ps = a.side + row * a.ld;
        for (int c = lane; c < d; c += 32) { const float s = ps[c]; sb[c] = s; tb[c] = pe[c] * s; }
        __syncwarp();
        float z[2] = {0.f, 0.f};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = lane + 32 * h;
            if (c < d) {
                float acc = bs[c];
                for (int k = 0; k < d; ++k) acc = fmaf(sb[k], Wg[k * d + c], fmaf(tb[k], Wb[k * d + c], acc));
                acc = acc > 0.f ? acc : 0.2f * acc;                                  // leaky_relu(., 0.2)
                if (!ngcf_keep(a.seed, a.step, a.layer, row, c, a.p_drop)) acc = 0.f; else acc *= keep_scale;
                z[h] = acc;
            }
        }
        float ss = z[0] * z[0] + z[1] * z[1];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const float nrm = fmaxf(sqrtf(ss), 1e-12f);                                  // F.normalize eps
        if (lane == 0) a.nrm[row] = nrm;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = lane + 32 * h;
            if (c < d) {
                a.ego_next[row * a.ld + c] = z[h];
                float *pa = a.acc + row * a.ld + c;
                *pa = *pa + a.acc_scale * (z[h] / nrm);
            }
        }
        __syncwarp();
    }
}

struct NgcfBwd {
    const float *gout;            // [N, ld]  dL/d(out) (NOT yet divided by L+1)
    const float *gnext;           // [N, ld]  gradient reaching ego_next from the layer above, or NULL (top layer)
    const float *ego, *side, *ego_next, *nrm;
    const float *Wg, *Wb;
    float *gz;                    // [N, ld]  out: gradient at the pre-activation
    float *gside;                 // [N, ld]  out: d(side)
    float *gego;                  // [N, ld]  out: direct part of d(ego)  ( + A_hat gside follows )
    float acc_scale;
    int N, ld, d, layer;
    float p_drop; uint64_t seed, step;
};

__global__ void __launch_bounds__(256) ngcf_bwd_row_kernel(const NgcfBwd a) {
    extern __shared__ float sm[];
    const int d = a.d;
    float *WgT = sm, *WbT = sm + d * d, *rowbuf = WbT + d * d;                      // transposed: [out][in]
    for (int e = threadIdx.x; e < d * d; e += blockDim.x) {
        const int k = e / d, c = e % d;
        WgT[c * d + k] = a.Wg[e]; WbT[c * d + k] = a.Wb[e];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float *gb = rowbuf + wid * d;
    const float keep_scale = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
    for (int64_t row = (int64_t)blockIdx.x * 8 + wid; row < a.N; row += (int64_t)gridDim.x * 8) {
        const float nrm = a.nrm[row];
        float e[2] = {0.f, 0.f}, gn[2] = {0.f, 0.f}, gx[2] = {0.f, 0.f};
        float dot = 0.f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = lane + 32 * h;
            if (c < d) {
                e[h] = a.ego_next[row * a.ld + c];
                gn[h] = a.gout[row * a.ld + c] * a.acc_scale;
                gx[h] = a.gnext ? a.gnext[row * a.ld + c] : 0.f;
                dot += (e[h] / nrm) * gn[h];
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = lane + 32 * h;
            if (c < d) {
                // normalize backward (norm above its eps), then what the next layer sends down, dropout, leaky relu
                float g = (gn[h] - (e[h] / nrm) * dot) / nrm + gx[h];
                if (!ngcf_keep(a.seed, a.step, a.layer, row, c, a.p_drop)) g = 0.f; else g *= keep_scale;
                // sign of the pre-activation = sign of the kept activation; a dropped element has zero gradient anyway
                g *= (e[h] > 0.f) ? 1.f : 0.2f;
                a.gz[row * a.ld + c] = g;
                gb[c] = g;
            }
        }
        __syncwarp();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = lane + 32 * h;
            if (k < d) {
                float gs = 0.f, gt = 0.f;
                for (int c = 0; c < d; ++c) { gs = fmaf(gb[c], WgT[c * d + k], gs); gt = fmaf(gb[c], WbT[c * d + k], gt); }
                const float ek = a.ego[row * a.ld + k], sk = a.side[row * a.ld + k];
                a.gside[row * a.ld + k] = gs + gt * ek;
                a.gego[row * a.ld + k] = gt * sk;
            }
        }
        __syncwarp();
    }
}

// dWg[k][c] += sum_rows side[k] gz[c];  dWb[k][c] += sum_rows (ego*side)[k] gz[c];  db[c] += sum_rows gz[c]
// CTA = 256 threads, each owns a (d/16) x (d/16) patch of both matrices; rows staged 32 at a time in shared memory.
__global__ void __launch_bounds__(256) ngcf_wgrad_kernel(const float *__restrict__ ego, const float *__restrict__ side,
                                                         const float *__restrict__ gz, int N, int ld, int d,
                                                         int rows_per_cta, float *dWg, float *dWb, float *db) {
    constexpr int TR = 32;
    __shared__ float s_side[TR][kNgcfMaxD + 1], s_t[TR][kNgcfMaxD + 1], s_g[TR][kNgcfMaxD + 1];
    const int tk = threadIdx.x / 16, tc = threadIdx.x % 16;       // patch rows k = tk + 16*a, cols c = tc + 16*b
    float accg[4][4], accb[4][4], accs[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) { accs[x] = 0.f;
#pragma unroll
        for (int y = 0; y < 4; ++y) accg[x][y] = accb[x][y] = 0.f; }
    const int64_t r_begin = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r_end = (r_begin + rows_per_cta < N) ? r_begin + rows_per_cta : N;
    for (int64_t r0 = r_begin; r0 < r_end; r0 += TR) {
        for (int e = threadIdx.x; e < TR * d; e += 256) {
            const int rr = e / d, c = e % d;
            const int64_t row = r0 + rr;
            float s = 0.f, t = 0.f, g = 0.f;
            if (row < r_end) { s = side[row * ld + c]; t = ego[row * ld + c] * s; g = gz[row * ld + c]; }
            s_side[rr][c] = s; s_t[rr][c] = t; s_g[rr][c] = g;
        }
        __syncthreads();
        for (int rr = 0; rr < TR; ++rr) {
            float sv[4], tv[4], gv[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const int k = tk + 16 * x, c = tc + 16 * x;
                sv[x] = k < d ? s_side[rr][k] : 0.f; tv[x] = k < d ? s_t[rr][k] : 0.f; gv[x] = c < d ? s_g[rr][c] : 0.f;
            }
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) { accg[x][y] = fmaf(sv[x], gv[y], accg[x][y]); accb[x][y] = fmaf(tv[x], gv[y], accb[x][y]); }
            if (tk == 0) {
#pragma unroll
                for (int y = 0; y < 4; ++y) accs[y] += gv[y];
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            const int k = tk + 16 * x, c = tc + 16 * y;
            if (k < d && c < d) { atomicAdd(dWg + k * d + c, accg[x][y]); atomicAdd(dWb + k * d + c, accb[x][y]); }
        }
    if (tk == 0)
#pragma unroll
        for (int y = 0; y < 4; ++y) { const int c = tc + 16 * y; if (c < d) atomicAdd(db + c, accs[y]); }
}

}  // namespace b200

using namespace b200;

extern "C" int b200rec_ngcf_layer_forward(const float *ego, const float *side, const float *W_gc, const float *b_gc,
                                          const float *W_bi, const float *b_bi, int n_rows, int ld, int d, int layer,
                                          float mess_dropout, uint64_t seed, uint64_t step, float *ego_next, float *nrm,
                                          float *acc, float acc_scale, void *stream) {
    B200_REQUIRE(ego && side && W_gc && b_gc && W_bi && b_bi && ego_next && nrm && acc, B200REC_EINVAL,
                 "ngcf_layer_forward: null argument");
    B200_REQUIRE(d >= 1 && d <= kNgcfMaxD && ld >= d, B200REC_EUNSUPPORTED, "ngcf_layer_forward: need 1 <= d <= %d", kNgcfMaxD);
    B200_REQUIRE(mess_dropout >= 0.f && mess_dropout < 1.f, B200REC_EINVAL, "ngcf_layer_forward: dropout in [0,1)");
    if (n_rows <= 0) return B200REC_OK;
    NgcfFwd a;
    a.ego = ego; a.side = side; a.Wg = W_gc; a.bg = b_gc; a.Wb = W_bi; a.bb = b_bi; a.ego_next = ego_next; a.nrm = nrm;
    a.acc = acc; a.acc_scale = acc_scale; a.N = n_rows; a.ld = ld; a.d = d; a.layer = layer; a.p_drop = mess_dropout;
    a.seed = seed; a.step = step;
    const size_t smem = sizeof(float) * ((size_t)2 * d * d + d + (size_t)8 * 2 * d);
    B200_CUDA(cudaFuncSetAttribute(ngcf_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = ((int64_t)n_rows + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 4;
    ngcf_fwd_kernel<<<(int)(blocks < cap ? blocks : cap), 256, smem, (cudaStream_t)stream>>>(a);
    B200_LAUNCH_CHECK();
    return B200REC_OK;
}

extern "C" int b200rec_ngcf_layer_backward(const float *g_out, const float *g_next, const float *ego, const float *side,
                                           const float *ego_next, const float *nrm, const float *W_gc, const float *W_bi,
                                           int n_rows, int ld, int d, int layer, float mess_dropout, uint64_t seed,
                                           uint64_t step, float acc_scale, float *g_z, float *g_side, float *g_ego,
                                           float *dW_gc, float *dW_bi, float *db, void *stream) {
    B200_REQUIRE(g_out && ego && side && ego_next && nrm && W_gc && W_bi && g_z && g_side && g_ego && dW_gc && dW_bi && db,
                 B200REC_EINVAL, "ngcf_layer_backward: null argument");
    B200_REQUIRE(d >= 1 && d <= kNgcfMaxD && ld >= d, B200REC_EUNSUPPORTED, "ngcf_layer_backward: need 1 <= d <= %d", kNgcfMaxD);
    if (n_rows <= 0) return B200REC_OK;
    cudaStream_t s = (cudaStream_t)stream;
    NgcfBwd a;
    a.gout = g_out; a.gnext = g_next; a.ego = ego; a.side = side; a.ego_next = ego_next; a.nrm = nrm; a.Wg = W_gc; a.Wb = W_bi;
    a.gz = g_z; a.gside = g_side; a.gego = g_ego; a.acc_scale = acc_scale; a.N = n_rows; a.ld = ld; a.d = d; a.layer = layer;
    a.p_drop = mess_dropout; a.seed = seed; a.step = step;
    const size_t smem = sizeof(float) * ((size_t)2 * d * d + (size_t)8 * d);
    B200_CUDA(cudaFuncSetAttribute(ngcf_bwd_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = ((int64_t)n_rows + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 4;
    ngcf_bwd_row_kernel<<<(int)(blocks < cap ? blocks : cap), 256, smem, s>>>(a);
    B200_LAUNCH_CHECK();
    int grid = sm_count() * 2;
    int rows_per_cta = (n_rows + grid - 1) / grid;
    rows_per_cta = (rows_per_cta + 31) / 32 * 32;
    grid = (n_rows + rows_per_cta - 1) / rows_per_cta;
    ngcf_wgrad_kernel<<<grid, 256, 0, s>>>(ego, side, g_z, n_rows, ld, d, rows_per_cta, dW_gc, dW_bi, db);
    B200_LAUNCH_CHECK();
    return B200REC_OK;
}
