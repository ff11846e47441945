// tvq_snake.cuh — the Snake activation of the stage-1 encoder / decoder stacks, y = x + sin^2(a_c x) / a_c with one
// learnable a per channel (/root/reference/timevqvae/utils/train_utils.py:421-448, a TorchScript module there), forward and
// backward as one kernel each.  NOT on the VQ hot path: it serves the stage-1 harness (stage1.py), where the eager torch
// expression (mul, sin, pow, reciprocal, mul, add; backward: eight more plus a full reduction per parameter) is a third of
// the step.  HBM-bound elementwise work: 8 bytes per element forward, 12 backward; the per-channel gradient of `a` is
// reduced in shared memory and leaves the CTA as one atomic per channel.
//   x viewed as [n, c, s]: channels_last = 0: element (i, j, t) at (i*c + j)*s + t (NCHW);  1: at (i*s + t)*c + j (NHWC)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tvq {

constexpr int kSnakeMaxC = 1024;

template <bool CL>
__device__ __forceinline__ int snake_channel(int64_t i, int c, int s) {
    return CL ? (int)(i % c) : (int)((i / s) % c);
}

template <bool CL>
__global__ void __launch_bounds__(256) snake_fwd_kernel(const float* __restrict__ x, const float* __restrict__ a, int64_t total,
                                                        int c, int s, float* __restrict__ y) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const float av = __ldg(a + snake_channel<CL>(i, c, s));
        const float xv = x[i];
        const float sn = sinf(av * xv);
        y[i] = xv + (1.0f / av) * (sn * sn);
    }
}

// g_x = g * (1 + sin(2 a x));   g_a[c] = sum g * (x sin(2 a x) / a - sin^2(a x) / a^2)
template <bool CL>
__global__ void __launch_bounds__(256) snake_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                        const float* __restrict__ a, int64_t total, int c, int s,
                                                        float* __restrict__ gx, float* __restrict__ ga) {
    extern __shared__ float acc[];                       // [c]
    for (int j = threadIdx.x; j < c; j += blockDim.x) acc[j] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    // warp-uniform trip count (the last trip may have idle lanes): the warp-level reduction below needs all 32 lanes
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < total; base += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = base + lane;
        const bool valid = i < total;
        const int ch = valid ? snake_channel<CL>(i, c, s) : -1;
        float part = 0.f;
        if (valid) {
            const float av = __ldg(a + ch);
            const float xv = x[i], gv = g[i];
            float sn, cs;
            sincosf(av * xv, &sn, &cs);
            const float s2 = 2.0f * sn * cs, inv = 1.0f / av;
            gx[i] = gv * (1.0f + s2);
            part = gv * (xv * s2 * inv - sn * sn * inv * inv);
        }
        if (!CL) {
            // NCHW: the 32 elements of a warp are (almost always) in one plane: one shared-memory atomic per warp
            const int ch0 = __shfl_sync(0xffffffffu, ch, 0);
            if (__all_sync(0xffffffffu, ch == ch0 || !valid)) {
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
                if (lane == 0) atomicAdd(acc + ch0, part);
            } else if (valid) {
                atomicAdd(acc + ch, part);
            }
        } else if (valid) {
            atomicAdd(acc + ch, part);
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < c; j += blockDim.x) if (acc[j] != 0.f) atomicAdd(ga + j, acc[j]);
}

}  // namespace tvq
