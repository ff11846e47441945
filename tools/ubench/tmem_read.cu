// Micro-benchmark (experiment, not product): tcgen05.ld throughput per SM with 1/4/8/16 reading warps.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__global__ void k(int iters, int unroll2, unsigned long long* out, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        uint32_t a[32], b[32];
        ld32(base + ((i * 64) & 511), a);
        if (unroll2) ld32(base + ((i * 64 + 32) & 511), b);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= a[j];
        if (unroll2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) acc ^= b[j];
        }
    }
    long long t1 = clock64();
    __syncthreads();
    if (lane == 0) out[warp] = (unsigned long long)(t1 - t0);
    sink[threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}
int main() {
    unsigned long long* out; uint32_t* sink;
    cudaMalloc(&out, 64 * 8); cudaMalloc(&sink, 1024 * 4);
    const int iters = 20000;
    for (int u2 = 0; u2 < 2; ++u2)
        for (int warps : {1, 4, 8, 16}) {
            k<<<1, warps * 32>>>(iters, u2, out, sink);
            cudaError_t e = cudaDeviceSynchronize();
            unsigned long long h[64];
            cudaMemcpy(h, out, 64 * 8, cudaMemcpyDeviceToHost);
            unsigned long long mx = 0;
            for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
            double bytes = (double)iters * (u2 ? 2 : 1) * 4096.0 * warps;
            printf("warps=%2d loads/iter=%d: %llu clk, %.1f B/clk/SM, %.1f B/clk/warp (%s)\n", warps, u2 + 1, mx, bytes / mx, bytes / mx / warps,
                   cudaGetErrorString(e));
        }
    return 0;
}
