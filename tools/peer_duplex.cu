// Two GPUs gather (and write back) random 512-byte rows from EACH OTHER at the same time: what the P2P-fused BPR step
// asks of NVLink.  Sweeps the peer table size (TLB reach) and sorted vs random row order.
//   nvcc -O3 -arch=sm_100a tools/peer_duplex.cu -o gpurun_build/peer_duplex ; needs 2 GPUs
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int U, bool WRITE>
__global__ void __launch_bounds__(256) k(float *T, const int *ids, int n, float *sink) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float acc = 0.f;
    for (int64_t b = w * U; b < n; b += nw * U) {
        float4 v[U]; int id[U];
#pragma unroll
        for (int u = 0; u < U; ++u) id[u] = (b + u < n) ? ids[b + u] : -1;
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = id[u] >= 0 ? *reinterpret_cast<const float4 *>(T + (int64_t)id[u] * 128 + lane * 4) : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            acc += v[u].x + v[u].y + v[u].z + v[u].w;
            if (WRITE && id[u] >= 0) { float4 x = v[u]; x.x += 1e-6f; *reinterpret_cast<float4 *>(T + (int64_t)id[u] * 128 + lane * 4) = x; }
        }
    }
    if (acc == 12345.678f) *sink = acc;
}

int main(int argc, char **argv) {
    const int n = 1000000;
    for (int64_t rows : {1000000LL, 4000000LL, 10000000LL}) {
        for (int sorted = 0; sorted < 2; ++sorted) {
            float *T[2], *sink[2]; int *ids[2]; cudaStream_t st[2]; cudaEvent_t e0[2], e1[2];
            std::vector<int> h(rows);
            for (int64_t i = 0; i < rows; ++i) h[i] = (int)i;
            std::mt19937 rng(1); std::shuffle(h.begin(), h.end(), rng);
            if (sorted) std::sort(h.begin(), h.begin() + n);
            for (int g = 0; g < 2; ++g) {
                CK(cudaSetDevice(g)); CK(cudaMalloc(&T[g], rows * 512)); CK(cudaMemset(T[g], 0, rows * 512));
                CK(cudaMalloc(&sink[g], 4)); CK(cudaMalloc(&ids[g], n * 4)); CK(cudaMemcpy(ids[g], h.data(), n * 4, cudaMemcpyHostToDevice));
                CK(cudaStreamCreate(&st[g])); CK(cudaEventCreate(&e0[g])); CK(cudaEventCreate(&e1[g]));
                cudaError_t e = cudaDeviceEnablePeerAccess(1 - g, 0); if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e); cudaGetLastError();
            }
            for (int g = 0; g < 2; ++g) { CK(cudaSetDevice(g)); CK(cudaDeviceSynchronize()); }
            auto run = [&](const char *nm, int which, bool write, bool duplex) {
                for (int rep = 0; rep < 2; ++rep) {
                    for (int g = 0; g < 2; ++g) {
                        if (!duplex && g != 0) continue;
                        CK(cudaSetDevice(g)); CK(cudaEventRecord(e0[g], st[g]));
                        for (int i = 0; i < 5; ++i) {
                            if (write) k<4, true><<<148 * 6, 256, 0, st[g]>>>(T[1 - g], ids[g], n, sink[g]);
                            else k<4, false><<<148 * 6, 256, 0, st[g]>>>(T[1 - g], ids[g], n, sink[g]);
                        }
                        CK(cudaEventRecord(e1[g], st[g]));
                    }
                    for (int g = 0; g < 2; ++g) { CK(cudaSetDevice(g)); CK(cudaDeviceSynchronize()); }
                }
                float ms; CK(cudaSetDevice(0)); CK(cudaEventElapsedTime(&ms, e0[0], e1[0])); ms /= 5;
                printf("peer table %5.1f GB  %-7s %-22s %.4f ms  %.0f M rows/s per GPU\n", rows * 512 / 1e9, sorted ? "sorted" : "random", nm, ms, n / ms / 1e3);
            };
            run("read  one-sided", 0, false, false);
            run("r+w   one-sided", 0, true, false);
            run("read  duplex", 0, false, true);
            run("r+w   duplex", 0, true, true);
            for (int g = 0; g < 2; ++g) { CK(cudaSetDevice(g)); cudaFree(T[g]); cudaFree(ids[g]); cudaFree(sink[g]); }
        }
    }
    return 0;
}
