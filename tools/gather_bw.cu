// Random 512-byte row gather (+ optional write-back) bandwidth on one GPU: what is the roofline of the BPR step's
// access pattern, as a function of rows in flight per warp and occupancy?   nvcc -O3 -arch=sm_100a tools/gather_bw.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// G lanes per row (row = 128 floats = 32 float4; each lane holds 32/G float4), U rows in flight per group
template <int G, int U, bool WRITE>
__global__ void __launch_bounds__(256) gather_kernel(float *T, const int *ids, int n, float *sink) {
    constexpr int CPL = 32 / G, GPW = 32 / G;
    const int lane = threadIdx.x & 31, sl = lane % G, sg = lane / G;
    const int64_t group = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * GPW + sg;
    const int64_t n_groups = (((int64_t)gridDim.x * blockDim.x) >> 5) * GPW;
    float acc = 0.f;
    for (int64_t b = group * U; b < n; b += n_groups * U) {
        float4 v[U][CPL];
        int id[U];
#pragma unroll
        for (int u = 0; u < U; ++u) id[u] = (b + u < n) ? ids[b + u] : -1;
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int k = 0; k < CPL; ++k)
                if (id[u] >= 0) v[u][k] = *reinterpret_cast<const float4 *>(T + (int64_t)id[u] * 128 + (sl + k * G) * 4);
                else v[u][k] = make_float4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                acc += v[u][k].x + v[u][k].y + v[u][k].z + v[u][k].w;
                if (WRITE && id[u] >= 0) {
                    float4 w = v[u][k]; w.x += 1e-6f;
                    *reinterpret_cast<float4 *>(T + (int64_t)id[u] * 128 + (sl + k * G) * 4) = w;
                }
            }
    }
    if (acc == 12345.678f) *sink = acc;
}

template <int G, int U, bool WRITE>
void run(const char *name, float *T, const int *ids, int n, float *sink, int ctas_per_sm) {
    int grid = 148 * ctas_per_sm;
    for (int i = 0; i < 3; ++i) gather_kernel<G, U, WRITE><<<grid, 256>>>(T, ids, n, sink);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    const int reps = 10;
    for (int i = 0; i < reps; ++i) gather_kernel<G, U, WRITE><<<grid, 256>>>(T, ids, n, sink);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    double bytes = (double)n * 512 * (WRITE ? 2 : 1);
    printf("%-28s G=%2d U=%2d ctas/SM=%d  %.4f ms  %.2f TB/s\n", name, G, U, ctas_per_sm, ms, bytes / ms / 1e9);
}

int main(int argc, char **argv) {
    const int64_t rows = argc > 1 ? atoll(argv[1]) : 1000000;
    const int n = argc > 2 ? atoi(argv[2]) : 1000000;
    float *T, *sink; int *ids;
    const bool peer = argc > 3 && atoi(argv[3]) == 1;     // table on GPU 1, kernels on GPU 0: NVLink peer gather
    if (peer) {
        CK(cudaSetDevice(1)); CK(cudaMalloc(&T, rows * 512)); CK(cudaMemset(T, 0, rows * 512)); CK(cudaDeviceSynchronize());
        CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0));
        printf("PEER mode: table lives on GPU 1, gathered from GPU 0 over NVLink\n");
    } else {
        CK(cudaMalloc(&T, rows * 512)); CK(cudaMemset(T, 0, rows * 512));
    }
    CK(cudaMalloc(&sink, 4));
    std::vector<int> h(rows);
    for (int64_t i = 0; i < rows; ++i) h[i] = (int)i;
    std::mt19937 rng(1); std::shuffle(h.begin(), h.end(), rng);
    CK(cudaMalloc(&ids, (size_t)n * 4)); CK(cudaMemcpy(ids, h.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
    printf("table %lld rows x 512 B = %.0f MB, %d unique random rows per launch\n", (long long)rows, rows * 512 / 1e6, n);
#define SWEEP(W, nm) \
    run<32, 1, W>(nm, T, ids, n, sink, 8); run<32, 2, W>(nm, T, ids, n, sink, 8); run<32, 4, W>(nm, T, ids, n, sink, 8); \
    run<32, 8, W>(nm, T, ids, n, sink, 8); run<32, 4, W>(nm, T, ids, n, sink, 3); run<32, 2, W>(nm, T, ids, n, sink, 3); \
    run<32, 8, W>(nm, T, ids, n, sink, 3); run<32, 16, W>(nm, T, ids, n, sink, 3); \
    run<8, 1, W>(nm, T, ids, n, sink, 8); run<8, 2, W>(nm, T, ids, n, sink, 8); run<8, 4, W>(nm, T, ids, n, sink, 4); \
    run<8, 2, W>(nm, T, ids, n, sink, 3); run<4, 1, W>(nm, T, ids, n, sink, 8); run<4, 2, W>(nm, T, ids, n, sink, 4);
    SWEEP(false, "read")
    SWEEP(true, "read+write")
    if (peer) {   // deeper: more rows in flight per warp / more warps
        run<32, 16, false>("read", T, ids, n, sink, 8); run<32, 8, false>("read", T, ids, n, sink, 6);
        run<8, 4, false>("read", T, ids, n, sink, 8); run<8, 8, false>("read", T, ids, n, sink, 4);
        run<32, 16, true>("read+write", T, ids, n, sink, 8); run<8, 4, true>("read+write", T, ids, n, sink, 8);
    }
    // sorted ids (sequential rows): the streaming bound of the same kernel
    std::sort(h.begin(), h.begin() + n);
    CK(cudaMemcpy(ids, h.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
    run<32, 4, false>("read sorted ids", T, ids, n, sink, 8);
    run<32, 4, true>("read+write sorted ids", T, ids, n, sink, 8);
    return 0;
}
