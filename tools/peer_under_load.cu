// Does a GPU that is busy saturating its own L2 / HBM still serve NVLink peer reads at full speed?
// GPU 0 gathers random 512-byte rows from GPU 1's table while GPU 1 (a) idles, (b) runs a local random
// gather + write-back over its own table, (c) additionally hammers a small table with vector atomics (the BPR
// step's item-row traffic).  nvcc -O3 -arch=sm_100a tools/peer_under_load.cu ; needs 2 GPUs.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void __launch_bounds__(256) gather(float *T, const int *ids, int n, float *sink, int write) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float acc = 0.f;
    for (int64_t b = w * 2; b < n; b += nw * 2) {
        float4 v[2]; int id[2];
        for (int u = 0; u < 2; ++u) id[u] = (b + u < n) ? ids[b + u] : -1;
        for (int u = 0; u < 2; ++u) v[u] = id[u] >= 0 ? *reinterpret_cast<const float4 *>(T + (int64_t)id[u] * 128 + lane * 4) : make_float4(0, 0, 0, 0);
        for (int u = 0; u < 2; ++u) {
            acc += v[u].x + v[u].y + v[u].z + v[u].w;
            if (write && id[u] >= 0) { float4 x = v[u]; x.x += 1e-6f; *reinterpret_cast<float4 *>(T + (int64_t)id[u] * 128 + lane * 4) = x; }
        }
    }
    if (acc == 12345.678f) *sink = acc;
}
// local load like the BPR step: user row read+write, two small-table rows read + vector-reduced
__global__ void __launch_bounds__(256) bprlike(float *U, float *V, const int *ids, int n, int ni, float *sink) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = w; b < n; b += nw) {
        const int u = ids[b];
        const unsigned h = (unsigned)u * 2654435761u;
        const int i = h % ni, j = (h >> 7) % ni;
        float4 a = *reinterpret_cast<const float4 *>(U + (int64_t)u * 128 + lane * 4);
        float4 x = *reinterpret_cast<const float4 *>(V + (int64_t)i * 128 + lane * 4);
        float4 y = *reinterpret_cast<const float4 *>(V + (int64_t)j * 128 + lane * 4);
        a.x += 1e-6f * (x.x - y.x);
        *reinterpret_cast<float4 *>(U + (int64_t)u * 128 + lane * 4) = a;
        asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(V + (int64_t)i * 128 + lane * 4), "f"(1e-9f), "f"(0.f), "f"(0.f), "f"(0.f) : "memory");
        asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(V + (int64_t)j * 128 + lane * 4), "f"(-1e-9f), "f"(0.f), "f"(0.f), "f"(0.f) : "memory");
    }
    if (sink == nullptr) U[0] = 0.f;
}

int main() {
    const int64_t rows = 1250000; const int n = 1000000, ni = 125000;
    float *T[2], *V[2], *sink[2]; int *ids[2]; cudaStream_t st[2]; cudaEvent_t e0, e1;
    std::vector<int> h(rows);
    for (int64_t i = 0; i < rows; ++i) h[i] = (int)i;
    std::mt19937 rng(1); std::shuffle(h.begin(), h.end(), rng);
    for (int g = 0; g < 2; ++g) {
        CK(cudaSetDevice(g)); CK(cudaMalloc(&T[g], rows * 512)); CK(cudaMemset(T[g], 0, rows * 512));
        CK(cudaMalloc(&V[g], (size_t)ni * 512)); CK(cudaMemset(V[g], 0, (size_t)ni * 512));
        CK(cudaMalloc(&sink[g], 4)); CK(cudaMalloc(&ids[g], n * 4)); CK(cudaMemcpy(ids[g], h.data(), n * 4, cudaMemcpyHostToDevice));
        CK(cudaStreamCreate(&st[g]));
        cudaError_t e = cudaDeviceEnablePeerAccess(1 - g, 0); (void)e; cudaGetLastError();
    }
    CK(cudaSetDevice(0)); CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int mode = 0; mode < 4; ++mode) {
        const char *nm[] = {"home idle", "home: local gather+write", "home: BPR-like local step", "home: BPR-like, reader also BPR-like locally"};
        for (int write = 0; write < 2; ++write) {
            for (int rep = 0; rep < 2; ++rep) {
                // keep the home GPU busy for the whole measurement
                CK(cudaSetDevice(1));
                if (mode == 1) for (int i = 0; i < 40; ++i) gather<<<148 * 6, 256, 0, st[1]>>>(T[1], ids[1], n, sink[1], 1);
                if (mode >= 2) for (int i = 0; i < 40; ++i) bprlike<<<148 * 3, 256, 0, st[1]>>>(T[1], V[1], ids[1], n, ni, sink[1]);
                CK(cudaSetDevice(0));
                cudaStream_t s2; CK(cudaStreamCreate(&s2));
                if (mode == 3) for (int i = 0; i < 40; ++i) bprlike<<<148 * 3, 256, 0, s2>>>(T[0], V[0], ids[0], n, ni, sink[0]);
                CK(cudaEventRecord(e0, st[0]));
                for (int i = 0; i < 5; ++i) gather<<<148 * 6, 256, 0, st[0]>>>(T[1], ids[0], n, sink[0], write);
                CK(cudaEventRecord(e1, st[0]));
                CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 5;
                for (int g = 0; g < 2; ++g) { CK(cudaSetDevice(g)); CK(cudaDeviceSynchronize()); }
                CK(cudaSetDevice(0)); CK(cudaStreamDestroy(s2));
                if (rep == 1) printf("%-46s peer %-10s %.3f ms per 1M rows  (%.0f M rows/s)\n", nm[mode], write ? "read+write" : "read", ms, n / ms / 1e3);
            }
        }
    }
    return 0;
}
