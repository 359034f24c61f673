// Microbenchmark: tcgen05.ld (TMEM -> registers) throughput per SM on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_bw tools/tmem_bw.cu && tools/tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int X>
__device__ __forceinline__ uint32_t ld_sum(uint32_t taddr);
template <>
__device__ __forceinline__ uint32_t ld_sum<32>(uint32_t taddr) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
        "%29,%30,%31}, [%32];\n\ttcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s ^= r[i];
    return s;
}

template <int X, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) bw_kernel(int iters, unsigned long long *cycles, uint32_t *sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 512; c += X) acc ^= ld_sum<X>(base + c);
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512));
}

template <int X, int WARPS>
void run(const char *name) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned long long *cyc; uint32_t *sink;
    cudaMalloc(&cyc, sms * 8); cudaMalloc(&sink, sms * WARPS * 32 * 4);
    const int iters = 2000;
    bw_kernel<X, WARPS><<<sms, WARPS * 32>>>(iters, cyc, sink);
    bw_kernel<X, WARPS><<<sms, WARPS * 32>>>(iters, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[256];
    cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; ++i) avg += (double)h[i]; avg /= sms;
    // bytes read per CTA: each warp reads 32 lanes x 512 cols x 4 B per iteration
    const double bytes = (double)iters * WARPS * 32 * 512 * 4;
    printf("%s: %s  %.1f cycles/iter  %.1f B/cycle/SM  (%d warps)\n", name, cudaGetErrorString(e), avg / iters, bytes / avg, WARPS);
    cudaFree(cyc); cudaFree(sink);
}

int main() {
    run<32, 4>("x32 4 warps");
    run<32, 8>("x32 8 warps");
    run<32, 16>("x32 16 warps");
    return 0;
}
