// One process, G GPUs, plain cudaDeviceEnablePeerAccess mappings (no IPC): every GPU gathers random 512-byte rows from
// ALL other GPUs' tables, either one peer after the other or round-robin across peers per row.  Separates "several
// peers at once are slow" from "IPC-mapped peers are slow" (the P2P step kernel sees the slowdown with IPC mappings).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
constexpr int MAXG = 8;
struct Peers { float *T[MAXG]; };

// mode 0: rows [0,n) -> peer (k / per) in order (one peer at a time);  mode 1: row k -> peer (k % np)
__global__ void __launch_bounds__(256) gather(Peers P, int np, const int *ids, int n, int mode, int write, float *sink) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int per = (n + np - 1) / np;
    float acc = 0.f;
    for (int64_t b = w * 2; b < n; b += nw * 2) {
        float4 v[2]; float *p[2];
        for (int u = 0; u < 2; ++u) {
            const int64_t k = b + u;
            p[u] = nullptr;
            if (k < n) { const int peer = mode ? (int)(k % np) : (int)(k / per); p[u] = P.T[peer] + (int64_t)ids[k] * 128 + lane * 4; }
        }
        for (int u = 0; u < 2; ++u) v[u] = p[u] ? *reinterpret_cast<const float4 *>(p[u]) : make_float4(0, 0, 0, 0);
        for (int u = 0; u < 2; ++u) {
            acc += v[u].x + v[u].y + v[u].z + v[u].w;
            if (write && p[u]) { float4 x = v[u]; x.x += 1e-6f; *reinterpret_cast<float4 *>(p[u]) = x; }
        }
    }
    if (acc == 12345.678f) *sink = acc;
}

int main() {
    int G = 0; CK(cudaGetDeviceCount(&G)); if (G > MAXG) G = MAXG;
    const int64_t rows = 1250000; const int n = 875000;
    float *T[MAXG], *sink[MAXG]; int *ids[MAXG]; cudaStream_t st[MAXG]; cudaEvent_t e0[MAXG], e1[MAXG];
    std::vector<int> h(rows);
    for (int64_t i = 0; i < rows; ++i) h[i] = (int)i;
    for (int g = 0; g < G; ++g) {
        std::mt19937 rng(g + 1); std::shuffle(h.begin(), h.end(), rng);
        CK(cudaSetDevice(g)); CK(cudaMalloc(&T[g], rows * 512)); CK(cudaMemset(T[g], 0, rows * 512));
        CK(cudaMalloc(&sink[g], 4)); CK(cudaMalloc(&ids[g], n * 4)); CK(cudaMemcpy(ids[g], h.data(), n * 4, cudaMemcpyHostToDevice));
        CK(cudaStreamCreate(&st[g])); CK(cudaEventCreate(&e0[g])); CK(cudaEventCreate(&e1[g]));
        for (int q = 0; q < G; ++q) if (q != g) { cudaDeviceEnablePeerAccess(q, 0); cudaGetLastError(); }
    }
    for (int g = 0; g < G; ++g) { CK(cudaSetDevice(g)); CK(cudaDeviceSynchronize()); }
    printf("%d GPUs, each gathers %d random 512-B rows spread over its %d peers (1.25M-row tables)\n", G, n, G - 1);
    for (int all = 0; all < 2; ++all)
        for (int write = 0; write < 2; ++write)
            for (int mode = 0; mode < 2; ++mode) {
                float worst = 0.f;
                for (int rep = 0; rep < 2; ++rep) {
                    for (int g = 0; g < (all ? G : 1); ++g) {
                        Peers P; int np = 0;
                        for (int q = 0; q < G; ++q) if (q != g) P.T[np++] = T[q];
                        CK(cudaSetDevice(g)); CK(cudaEventRecord(e0[g], st[g]));
                        for (int i = 0; i < 4; ++i) gather<<<148 * 6, 256, 0, st[g]>>>(P, np, ids[g], n, mode, write, sink[g]);
                        CK(cudaEventRecord(e1[g], st[g]));
                    }
                    worst = 0.f;
                    for (int g = 0; g < (all ? G : 1); ++g) {
                        CK(cudaSetDevice(g)); CK(cudaDeviceSynchronize());
                        float ms; CK(cudaEventElapsedTime(&ms, e0[g], e1[g])); ms /= 4; if (ms > worst) worst = ms;
                    }
                }
                printf("%-18s %-10s %-22s %.3f ms  (%.0f M rows/s per GPU)\n", all ? "all GPUs at once" : "GPU 0 only",
                       write ? "read+write" : "read", mode ? "round-robin over peers" : "one peer at a time", worst, n / worst / 1e3);
            }
    return 0;
}
