// tvq_common.cuh — shared device helpers of the B200 VQ kernels (sm_100a only).
//
// Everything the scoring paths (SIMT and tcgen05) have in common lives here:
//  * the shared-memory tile layout (TMA/UMMA SWIZZLE_128B, K-major) and its address helpers,
//  * the canonical fp64 re-score that makes the decision independent of the scoring precision
//    (mirrored bit for bit by oracle/vq_canon.c),
//  * small PTX wrappers (cp.async, vector red, streaming stores).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tvq {

constexpr int kBM = 128;        // latents (rows) per CTA tile
constexpr int kThreads = 256;   // threads per CTA of the forward kernels
constexpr int kWarps = kThreads / 32;

// Scratch header at the start of the caller-provided workspace (zeroed once by the host,
// loss/diagnostics re-zeroed by the prep kernel of every call, ticket self-resetting).
struct WsHeader {
    double loss_sum;       // sum over all latents of (q_st - x)^2
    unsigned ticket;       // CTAs finished (last one computes the scalars)
    unsigned n_rescored;   // rows decided by the fp64 re-score
    unsigned n_exact;      // rows that needed the full exact scan
    unsigned ema_ticket;   // same, for the EMA kernel
    unsigned next_tile;    // dynamic tile scheduler of the resident-codebook forward (reset by the last CTA)
    unsigned pad[9];
};
static_assert(sizeof(WsHeader) == 64, "WsHeader is 64 bytes");

// ---------------------------------------------------------------------------------------------
// Tile layout.  A tile of ROWS rows x DP floats is stored as DP/32 sub-tiles; each sub-tile is
// ROWS rows x 128 bytes with the 16-byte chunk index XOR-ed with (row & 7): exactly what a TMA
// box of {32 floats, ROWS rows} with CU_TENSOR_MAP_SWIZZLE_128B writes and what a K-major
// SWIZZLE_128B UMMA descriptor reads.  The SIMT path uses the same layout so that every phase
// after scoring is shared; its loads of 4 consecutive rows are bank-conflict free.
// Returns the float index of 16-byte chunk `c4` of `row`.
template <int ROWS>
__device__ __forceinline__ int tile_off(int row, int c4) {
    return (c4 >> 3) * (ROWS * 32) + row * 32 + (((c4 & 7) ^ (row & 7)) << 2);
}

// ---------------------------------------------------------------------------------------------
// PTX wrappers
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// 16-byte reduction to global memory (no return value): one L2 atomic op per 4 floats.
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}
// Streaming (evict-first) 16-byte store for outputs that are written once.
__device__ __forceinline__ void st_stream_v4(float* addr, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ float4 ld_stream_v4(const float* addr) {
    float4 v;
    asm volatile("ld.global.cs.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(addr));
    return v;
}
__device__ __forceinline__ float ld_stream_v1(const float* addr) {
    float v;
    asm volatile("ld.global.cs.f32 %0, [%1];" : "=f"(v) : "l"(addr));
    return v;
}
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// ---------------------------------------------------------------------------------------------
// Canonical fp64 reduction tree (oracle/vq_canon.c::canon_dot): lane l owns the 16-byte chunks
// l, l+32, ... and adds their elements in increasing order; lanes are combined by an xor
// butterfly 16,8,4,2,1.  Every lane returns the same value.
__device__ __forceinline__ double butterfly_sum(double p) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) p += __shfl_xor_sync(0xffffffffu, p, off);
    return p;
}
__device__ __forceinline__ double dot4(double acc, float4 a, float4 b) {
    acc = fma((double)a.x, (double)b.x, acc);
    acc = fma((double)a.y, (double)b.y, acc);
    acc = fma((double)a.z, (double)b.z, acc);
    acc = fma((double)a.w, (double)b.w, acc);
    return acc;
}
// Canonical score d_k = fl32(fl32(x2 - fl32(2 x.e)) + e2)  (the reference's -dist, vq.py:210-214).
__device__ __forceinline__ float canon_score(float x2, double xe, float e2) {
    float xe2 = __double2float_rn(2.0 * xe);
    return __fadd_rn(__fsub_rn(x2, xe2), e2);
}
// Warp-cooperative canonical |e|^2 or x.e for rows in GLOBAL memory (d % 4 == 0, 16-byte aligned).
__device__ __forceinline__ double canon_dot_global(const float* a, const float* b, int nchunk, int lane) {
    double p = 0.0;
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    for (int c = lane; c < nchunk; c += 32) p = dot4(p, __ldg(a4 + c), __ldg(b4 + c));
    return butterfly_sum(p);
}

// Data-parallel exchange: spin until the local flag *lf shows `epoch` (written by a peer over NVLink with st.release.sys).
// A peer that does not show up within timeout_ns (0 = wait for ever) is REPORTED, not fatal: the step counter is stored in
// *err (the error word of this rank's exchange buffer, read by the host: PeerExchange.check()) and the wait gives up, so
// the CUDA context survives — what NCCL's watchdog does with a lost rank is the host's decision here too.
__device__ __forceinline__ void wait_flag_sys(const unsigned* lf, unsigned epoch, unsigned long long timeout_ns, unsigned* err) {
    unsigned long long t0 = 0;
    for (unsigned spin = 1;; ++spin) {
        unsigned v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(lf) : "memory");
        if (v == epoch) return;
        if ((spin & 0x3ffu) == 0 && timeout_ns != 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            else if (t - t0 > timeout_ns) { atomicMax(err, epoch); return; }
        }
    }
}

// Block-wide sum of one double per thread (kThreads threads); result valid in thread 0.
__device__ __forceinline__ double block_sum(double v, double* scratch /* >= kWarps doubles */) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) {
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += scratch[w];
    }
    return t;
}

}  // namespace tvq
