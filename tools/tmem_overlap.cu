// Microbenchmark: does ALU work overlap an in-flight tcgen05.ld in the same warp?  (sm_100a)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_overlap tools/tmem_overlap.cu && tools/tmem_overlap
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define LD64_OUT(r) "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), \
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), \
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), \
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), \
          "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), \
          "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), \
          "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), \
          "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
#define PIN32(r, o) "+r"(r[o+0]), "+r"(r[o+1]), "+r"(r[o+2]), "+r"(r[o+3]), "+r"(r[o+4]), "+r"(r[o+5]), "+r"(r[o+6]), "+r"(r[o+7]), \
          "+r"(r[o+8]), "+r"(r[o+9]), "+r"(r[o+10]), "+r"(r[o+11]), "+r"(r[o+12]), "+r"(r[o+13]), "+r"(r[o+14]), "+r"(r[o+15]), \
          "+r"(r[o+16]), "+r"(r[o+17]), "+r"(r[o+18]), "+r"(r[o+19]), "+r"(r[o+20]), "+r"(r[o+21]), "+r"(r[o+22]), "+r"(r[o+23]), \
          "+r"(r[o+24]), "+r"(r[o+25]), "+r"(r[o+26]), "+r"(r[o+27]), "+r"(r[o+28]), "+r"(r[o+29]), "+r"(r[o+30]), "+r"(r[o+31])

__device__ __forceinline__ void ld64(uint32_t taddr, uint32_t (&r)[64]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
        "%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,"
        "%56,%57,%58,%59,%60,%61,%62,%63}, [%64];" : LD64_OUT(r) : "r"(taddr));
}
__device__ __forceinline__ void wait64(uint32_t (&r)[64]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : PIN32(r, 0) :: "memory");
    asm volatile("" : PIN32(r, 32) :: "memory");
}
__device__ __forceinline__ float max3(float a, float b, float c) {
    float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;
}
__device__ __forceinline__ float tree(const uint32_t (&r)[64]) {
    float gm[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        const float a = max3(__uint_as_float(r[8 * g]), __uint_as_float(r[8 * g + 1]), __uint_as_float(r[8 * g + 2]));
        const float b = max3(__uint_as_float(r[8 * g + 3]), __uint_as_float(r[8 * g + 4]), __uint_as_float(r[8 * g + 5]));
        gm[g] = max3(a, b, fmaxf(__uint_as_float(r[8 * g + 6]), __uint_as_float(r[8 * g + 7])));
    }
    return max3(max3(gm[0], gm[1], gm[2]), max3(gm[3], gm[4], gm[5]), fmaxf(gm[6], gm[7]));
}

// MODE 0: ld,wait   1: ld,wait,tree   2: pipelined (ld next | tree current | wait)   3: tree only (no ld in the loop)
template <int MODE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k(int iters, unsigned long long *cycles, float *sink) {
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
    float acc = -1e30f;
    uint32_t ra[64], rb[64];
    ld64(base, ra); wait64(ra);
    ld64(base + 64, rb); wait64(rb);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
            ld64(base + 128, ra); wait64(ra);
            ld64(base + 192, rb); wait64(rb);
        } else if (MODE == 1) {
            ld64(base + 128, ra); wait64(ra); acc = fmaxf(acc, tree(ra));
            ld64(base + 192, rb); wait64(rb); acc = fmaxf(acc, tree(rb));
        } else if (MODE == 2) {
            ld64(base + 128, rb); acc = fmaxf(acc, tree(ra)); wait64(rb);
            ld64(base + 192, ra); acc = fmaxf(acc, tree(rb)); wait64(ra);
        } else {
            ra[0] ^= (uint32_t)it; acc = fmaxf(acc, tree(ra));
            rb[0] ^= (uint32_t)it; acc = fmaxf(acc, tree(rb));
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc + __uint_as_float(ra[5]) + __uint_as_float(rb[7]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512));
}

template <int MODE, int WARPS>
void run(const char *name) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned long long *cyc; float *sink;
    cudaMalloc(&cyc, sms * 8); cudaMalloc(&sink, sms * WARPS * 32 * 4);
    const int iters = 4000;
    k<MODE, WARPS><<<sms, WARPS * 32>>>(iters, cyc, sink);
    k<MODE, WARPS><<<sms, WARPS * 32>>>(iters, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[256];
    cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; ++i) avg += (double)h[i]; avg /= sms;
    printf("%-28s %d warps: %s  %.1f cycles per x64 chunk per warp\n", name, WARPS, cudaGetErrorString(e), avg / iters / 2);
    cudaFree(cyc); cudaFree(sink);
}

int main() {
    run<0, 4>("ld+wait"); run<1, 4>("ld+wait+tree"); run<2, 4>("pipelined ld|tree"); run<3, 4>("tree only (dyn idx)");
    run<0, 8>("ld+wait"); run<1, 8>("ld+wait+tree"); run<2, 8>("pipelined ld|tree");
    return 0;
}
