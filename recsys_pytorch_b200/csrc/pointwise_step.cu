// Fused POINTWISE matrix-factorisation step (sm_100a) - the 2-row sibling of bpr_step.cu.
//
// Replaces, for the reference's pointwise mode (hparams['pointwise'] = True), models/MF.py:63-68 with
// MF.py:101-102: `loss_func(forward(users, items), ratings)` -> backward -> optimiser, where loss_func is
// F.binary_cross_entropy_with_logits ('ce', MF.py:21) or F.mse_loss ('mse'), both with reduction='mean':
//     x_b = U[u_b] . V[i_b]
//     'ce' : loss_b = (1-y) x + max(-x,0) + log(exp(-max(-x,0)) + exp(-x-max(-x,0))),  g_b = (sigmoid(x_b) - y_b) / B
//     'mse': loss_b = (x - y)^2,                                                        g_b = 2 (x_b - y_b) / B
//     dU[u_b] += g_b V[i_b],  dV[i_b] += g_b U[u_b]       (+ reg/B * row: the engine's per-occurrence L2, reg=0 = reference)
// One sub-group of G lanes per sample (G*16 B >= row bytes for d <= 128), both rows gathered once, the next sample's
// rows in flight while the current one reduces; sinks as in bpr_step.cu: UPDATE (in place, vector atomics - a batch of the
// reference's PointwiseGenerator repeats users AND items, so both rows use REDG), GRAD (dense gradient buffers for the
// reference's dense Adam), NONE (loss only).
// Oracle: oracle/bpr_oracle.py::pointwise_loss / pointwise_grads, pinned to tests/golden/tiny_pointwise.npz.
#include <math.h>
#include "common.cuh"

namespace b200 {

struct PwParams {
    float *U, *V;
    int ld, B, loss_kind, sink;
    const int32_t *users, *items;
    const float *ratings;
    float lr, regB, invB;
    float *gU, *gV;
    double *loss_sum;
};

template <int G>
__device__ __forceinline__ float pw_group_sum(float v) {
#pragma unroll
    for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int G, int CPL>
__global__ void __launch_bounds__(256) pointwise_step_kernel(const PwParams p) {
    constexpr int SPW = 32 / G;                       // samples per warp pass
    const int lane = threadIdx.x & 31;
    const int sl = lane % G, sg = lane / G;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int d4 = p.ld >> 2;
    const int64_t ld = p.ld;
    float loss_local = 0.f;

    for (int64_t base = warp_global * SPW; base < p.B; base += n_warps * SPW) {
        const int64_t t = base + sg;
        const bool valid = t < p.B;                   // the whole sub-group agrees
        int u = 0, i = 0;
        float y = 0.f;
        if (valid) { u = p.users[t]; i = p.items[t]; y = p.ratings[t]; }
        float4 ru[CPL], ri[CPL];
        float part = 0.f;
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            const int q = sl + k * G;
            if (valid && q < d4) {
                ru[k] = ld4(p.U + (int64_t)u * ld + q * 4);
                ri[k] = ld4(p.V + (int64_t)i * ld + q * 4);
            } else {
                ru[k] = ri[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            part = fmaf(ru[k].x, ri[k].x, part); part = fmaf(ru[k].y, ri[k].y, part);
            part = fmaf(ru[k].z, ri[k].z, part); part = fmaf(ru[k].w, ri[k].w, part);
        }
        const float x = pw_group_sum<G>(part);
        float g, l;
        if (p.loss_kind == 1) {                       // F.mse_loss
            g = 2.f * (x - y) * p.invB;
            l = (x - y) * (x - y);
        } else {                                      // F.binary_cross_entropy_with_logits
            const float s = 1.f / (1.f + expf(-x));
            g = (s - y) * p.invB;
            const float mx = fmaxf(-x, 0.f);
            l = (1.f - y) * x + mx + logf(expf(-mx) + expf(-x - mx));
        }
        if (!valid) continue;
        if (sl == 0 && p.loss_sum) loss_local += l;
        if (p.sink == B200REC_SINK_NONE) continue;
        const float sc = (p.sink == B200REC_SINK_GRAD) ? 1.f : -p.lr;
        float *dstU = (p.sink == B200REC_SINK_GRAD) ? p.gU : p.U;
        float *dstV = (p.sink == B200REC_SINK_GRAD) ? p.gV : p.V;
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            const int q = sl + k * G;
            if (q < d4) {
                float4 du, di;
                du.x = sc * fmaf(g, ri[k].x, p.regB * ru[k].x); du.y = sc * fmaf(g, ri[k].y, p.regB * ru[k].y);
                du.z = sc * fmaf(g, ri[k].z, p.regB * ru[k].z); du.w = sc * fmaf(g, ri[k].w, p.regB * ru[k].w);
                di.x = sc * fmaf(g, ru[k].x, p.regB * ri[k].x); di.y = sc * fmaf(g, ru[k].y, p.regB * ri[k].y);
                di.z = sc * fmaf(g, ru[k].z, p.regB * ri[k].z); di.w = sc * fmaf(g, ru[k].w, p.regB * ri[k].w);
                red4(dstU + (int64_t)u * ld + q * 4, du);
                red4(dstV + (int64_t)i * ld + q * 4, di);
            }
        }
    }
    if (p.loss_sum) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) loss_local += __shfl_xor_sync(0xffffffffu, loss_local, o);
        if (lane == 0 && loss_local != 0.f) atomicAdd(p.loss_sum, (double)loss_local);
    }
}

template <int G, int CPL>
static int launch_pointwise(const PwParams &p, cudaStream_t s) {
    constexpr int SPW = 32 / G;
    int64_t blocks = ((int64_t)p.B + 8 * SPW - 1) / (8 * SPW);
    const int64_t cap = (int64_t)sm_count() * 8;
    pointwise_step_kernel<G, CPL><<<(int)(blocks < cap ? blocks : cap), 256, 0, s>>>(p);
    B200_LAUNCH_CHECK();
    return B200REC_OK;
}

}  // namespace b200

using namespace b200;

extern "C" int b200rec_pointwise_step(float *U, float *V, int ld, int d, const int32_t *users, const int32_t *items,
                                      const float *ratings, int B, int loss_kind, float lr, float reg, int sink,
                                      float *gU, float *gV, double *loss_sum, float inv_batch, void *stream) {
    B200_REQUIRE(U && V && users && items && ratings, B200REC_EINVAL, "pointwise_step: null argument");
    B200_REQUIRE(d >= 1 && ld >= d && ld % 4 == 0 && ld <= 512, B200REC_EINVAL, "pointwise_step: bad d/ld (%d/%d)", d, ld);
    B200_REQUIRE(loss_kind == 0 || loss_kind == 1, B200REC_EINVAL, "pointwise_step: loss_kind 0 (ce) or 1 (mse)");
    B200_REQUIRE(sink == B200REC_SINK_UPDATE || sink == B200REC_SINK_GRAD || sink == B200REC_SINK_NONE, B200REC_EINVAL,
                 "pointwise_step: sink must be UPDATE, GRAD or NONE");
    B200_REQUIRE(sink != B200REC_SINK_GRAD || (gU && gV), B200REC_EINVAL, "pointwise_step: SINK_GRAD needs gU, gV");
    if (B <= 0) return B200REC_OK;
    PwParams p;
    p.U = U; p.V = V; p.ld = ld; p.B = B; p.loss_kind = loss_kind; p.sink = sink;
    p.users = users; p.items = items; p.ratings = ratings;
    p.invB = inv_batch > 0.f ? inv_batch : 1.0f / (float)B;
    p.lr = lr; p.regB = reg * p.invB;
    p.gU = gU; p.gV = gV; p.loss_sum = loss_sum;
    const int d4 = ld / 4;
    int G = 1;
    while (G < d4 && G < 32) G <<= 1;
    const int CPL = (d4 + G - 1) / G;
    cudaStream_t s = (cudaStream_t)stream;
    switch (G) {
        case 1: return launch_pointwise<1, 1>(p, s);
        case 2: return launch_pointwise<2, 1>(p, s);
        case 4: return launch_pointwise<4, 1>(p, s);
        case 8: return launch_pointwise<8, 1>(p, s);
        case 16: return launch_pointwise<16, 1>(p, s);
        default:
            switch (CPL) {
                case 1: return launch_pointwise<32, 1>(p, s);
                case 2: return launch_pointwise<32, 2>(p, s);
                case 3: return launch_pointwise<32, 3>(p, s);
                default: return launch_pointwise<32, 4>(p, s);
            }
    }
}
