// tvq_aux.cuh — the small kernels around the fused forward: per-call preparation, EMA codebook
// update, backward, de-tokenising gather, dense distance matrix and dead-code re-seed.
// All of them are HBM- or latency-bound elementwise / gather work (no tensor cores).
#pragma once
#include <cuda_bf16.h>

#include "tvq_common.cuh"

namespace tvq {

// ------------------------------------------------------------------------------------------
// Per-call preparation: canonical |e_k|^2 (one warp per code), zero the statistics buffer and
// the loss / diagnostic fields of the workspace header; optionally the bf16 copy of the codebook
// ([k, dp] row-major, zero padded to dp columns) that the streamed tcgen05 path feeds to TMA.
__global__ void __launch_bounds__(256) prep_kernel(const float* __restrict__ cb, int k, int d, float* __restrict__ e2,
                                                   WsHeader* hdr, float* __restrict__ stats, int64_t stats_len,
                                                   __nv_bfloat16* __restrict__ cbh, int dp) {
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int c = gwarp; c < k; c += nwarps) {
        const float* er = cb + (size_t)c * d;
        double s = canon_dot_global(er, er, d >> 2, lane);
        if (lane == 0) e2[c] = __double2float_rn(s);
    }
    for (int c = k + gwarp * 32 + lane; c < ((k + 255) & ~255); c += nwarps * 32) e2[c] = 1e30f;   // pad: never the minimum
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gsz = (int64_t)gridDim.x * blockDim.x;
    if (stats != nullptr) {
        float4* s4 = reinterpret_cast<float4*>(stats);
        for (int64_t i = gtid; i < (stats_len >> 2); i += gsz) s4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int64_t i = (stats_len & ~int64_t(3)) + gtid; i < stats_len; i += gsz) stats[i] = 0.f;
    }
    if (cbh != nullptr) {
        const int dq = dp >> 2, nchunk = d >> 2;
        const int64_t total = (int64_t)k * dq;
        for (int64_t f = gtid; f < total; f += gsz) {
            const int64_t row = f / dq;
            const int c4 = (int)(f - row * dq);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c4 < nchunk) v = __ldg(reinterpret_cast<const float4*>(cb + (size_t)row * d) + c4);
            const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
            uint2 o;
            o.x = *reinterpret_cast<const uint32_t*>(&lo);
            o.y = *reinterpret_cast<const uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(cbh + (size_t)row * dp + 4 * c4) = o;
        }
    }
    if (gtid == 0) {
        hdr->loss_sum = 0.0;
        hdr->n_rescored = 0u;
        hdr->n_exact = 0u;
    }
}

// ------------------------------------------------------------------------------------------
// EMA codebook update, one launch (vq.py:231,236-242; ema_inplace :59-60, laplace_smoothing :63-64).
//   cs'      = cs * decay + counts * (1 - decay)
//   avg'     = avg * decay + embed_sum * (1 - decay)
//   n        = sum(cs');  smoothed = (cs' + eps) / (n + k*eps) * n;   embed = avg' / smoothed
// Every CTA derives n from the OLD cluster sizes (read-only during the kernel); the last CTA to
// finish writes cs' back, so no CTA ever reads a half-updated buffer.
struct EmaParams {
    const float* stats;
    float* cluster_size;
    float* embed_avg;
    float* embed;
    float* embed_prev;   // optional: receives the pre-update codebook (for the backward)
    int k, d;
    float decay, one_minus_decay, eps, k_eps;
    WsHeader* hdr;
};

__device__ __forceinline__ float ema_mix(float old_v, float new_v, float decay, float omd) {
    // mul_(decay) then add_(new, alpha=1-decay): the product is rounded, then alpha*new + that
    return fmaf(new_v, omd, __fmul_rn(old_v, decay));
}

__global__ void __launch_bounds__(256) ema_kernel(const EmaParams p) {
    __shared__ double red[8];
    __shared__ float s_n;
    __shared__ int s_last;
    const int tid = threadIdx.x;
    double part = 0.0;
    for (int c = tid; c < p.k; c += blockDim.x)
        part += (double)ema_mix(__ldcg(p.cluster_size + c), __ldg(p.stats + c), p.decay, p.one_minus_decay);
    double tot = block_sum(part, red);
    if (tid == 0) s_n = __double2float_rn(tot);
    __syncthreads();
    const float n = s_n;
    const float denom = __fadd_rn(n, p.k_eps);
    const int kp = (p.k + 3) & ~3;
    const int dq = p.d >> 2;
    const int64_t total = (int64_t)p.k * dq;
    const float4* esum4 = reinterpret_cast<const float4*>(p.stats + kp);
    float4* avg4 = reinterpret_cast<float4*>(p.embed_avg);
    float4* emb4 = reinterpret_cast<float4*>(p.embed);
    float4* prev4 = reinterpret_cast<float4*>(p.embed_prev);
    for (int64_t f = (int64_t)blockIdx.x * blockDim.x + tid; f < total; f += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(f / dq);
        const float cs = ema_mix(__ldcg(p.cluster_size + c), __ldg(p.stats + c), p.decay, p.one_minus_decay);
        const float sm = __fmul_rn(__fdiv_rn(__fadd_rn(cs, p.eps), denom), n);
        const float4 s = __ldg(esum4 + f);
        float4 a = avg4[f];
        a.x = ema_mix(a.x, s.x, p.decay, p.one_minus_decay);
        a.y = ema_mix(a.y, s.y, p.decay, p.one_minus_decay);
        a.z = ema_mix(a.z, s.z, p.decay, p.one_minus_decay);
        a.w = ema_mix(a.w, s.w, p.decay, p.one_minus_decay);
        avg4[f] = a;
        if (prev4) prev4[f] = emb4[f];
        emb4[f] = make_float4(__fdiv_rn(a.x, sm), __fdiv_rn(a.y, sm), __fdiv_rn(a.z, sm), __fdiv_rn(a.w, sm));
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        unsigned t = atomicAdd(&p.hdr->ema_ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        for (int c = tid; c < p.k; c += blockDim.x)
            p.cluster_size[c] = ema_mix(__ldcg(p.cluster_size + c), __ldg(p.stats + c), p.decay, p.one_minus_decay);
        if (tid == 0) p.hdr->ema_ticket = 0;
    }
}

// ------------------------------------------------------------------------------------------
// Backward: g_x = g_q + coef * (x - q_st), coef = g_loss * w * 2 / (n*d); q_st = x + (e[idx]-x).
// 12d + 8 bytes per latent (the code word comes from the L2-resident codebook).
__global__ void __launch_bounds__(256) backward_kernel(const float* __restrict__ g_q, const float* __restrict__ g_commit,
                                                        const float* __restrict__ g_weighted, const float* __restrict__ x,
                                                        const int64_t* __restrict__ idx, const float* __restrict__ cb, int64_t n,
                                                        int d, float weight, float scale, float* __restrict__ g_x) {
    const int dq = d >> 2;
    const int64_t total = n * dq;
    const float coef = fmaf(weight, g_weighted ? __ldg(g_weighted) : 0.f, g_commit ? __ldg(g_commit) : 0.f) * scale;
    for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < total; f += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = f / dq;
        const int c = (int)(f - row * dq);
        const int64_t code = __ldg(idx + row);
        const float4 xv = ld_stream_v4(x + 4 * f);
        const float4 gv = g_q ? ld_stream_v4(g_q + 4 * f) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 ev = __ldg(reinterpret_cast<const float4*>(cb + (size_t)code * d) + c);
        float4 o;
        o.x = fmaf(coef, __fsub_rn(xv.x, __fadd_rn(xv.x, __fsub_rn(ev.x, xv.x))), gv.x);
        o.y = fmaf(coef, __fsub_rn(xv.y, __fadd_rn(xv.y, __fsub_rn(ev.y, xv.y))), gv.y);
        o.z = fmaf(coef, __fsub_rn(xv.z, __fadd_rn(xv.z, __fsub_rn(ev.z, xv.z))), gv.z);
        o.w = fmaf(coef, __fsub_rn(xv.w, __fadd_rn(xv.w, __fsub_rn(ev.w, xv.w))), gv.w);
        st_stream_v4(g_x + 4 * f, o);
    }
}

// ------------------------------------------------------------------------------------------
// De-tokenising gather (models/maskgit.py:465-470).
// layout 0: out[b, t, :] = cb[tok[b, t]]       — one warp per token, 16-byte lanes
__global__ void __launch_bounds__(256) gather_rows_kernel(const int64_t* __restrict__ tok, const float* __restrict__ cb,
                                                           int64_t ntok, int k, int d, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int dq = d >> 2;
    for (int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < ntok; t += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        int64_t code = __ldg(tok + t);
        code = code < 0 ? 0 : (code >= k ? k - 1 : code);
        const float4* er = reinterpret_cast<const float4*>(cb + (size_t)code * d);
        for (int c = lane; c < dq; c += 32) st_stream_v4(out + (size_t)t * d + 4 * c, __ldg(er + c));
    }
}
// layout 1: out[b, :, t] = cb[tok[b, t]]  (decoder layout "b c (h w)"): 32 tokens x 32 channels
// are transposed through shared memory so that both the code-word reads and the output writes
// are coalesced.
__global__ void __launch_bounds__(256) gather_transposed_kernel(const int64_t* __restrict__ tok, const float* __restrict__ cb,
                                                                 int64_t b, int64_t t, int k, int d, float* __restrict__ out) {
    __shared__ float tile[32][33];
    __shared__ int codes[32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 8 warps
    const int64_t t_tiles = (t + 31) / 32, d_tiles = (d + 31) / 32;
    const int64_t total = b * t_tiles * d_tiles;
    for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
        const int64_t bi = w / (t_tiles * d_tiles);
        const int64_t rem = w - bi * t_tiles * d_tiles;
        const int64_t t0 = (rem / d_tiles) * 32, d0 = (rem % d_tiles) * 32;
        __syncthreads();
        if (threadIdx.x < 32) {
            int64_t tt = t0 + threadIdx.x;
            int64_t code = tt < t ? __ldg(tok + bi * t + tt) : 0;
            codes[threadIdx.x] = (int)(code < 0 ? 0 : (code >= k ? k - 1 : code));
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8)   // r: token within tile, tx: channel
            tile[r][tx] = (d0 + tx < d) ? __ldg(cb + (size_t)codes[r] * d + d0 + tx) : 0.f;
        __syncthreads();
        for (int r = ty; r < 32; r += 8)   // r: channel within tile, tx: token
            if (d0 + r < d && t0 + tx < t) out[(bi * d + d0 + r) * t + t0 + tx] = tile[tx][r];
    }
}

// ------------------------------------------------------------------------------------------
// Dense -dist matrix for the stochastic branch (vq.py:210-214, fp32 three-term formula).
// One warp per latent; k is small where this is used (stage 3 / sampler, k = 32).
__global__ void __launch_bounds__(256) neg_dist_kernel(const float* __restrict__ x, const float* __restrict__ cb,
                                                        int64_t n, int k, int d, float* __restrict__ dist) {
    const int lane = threadIdx.x & 31;
    const int dq = d >> 2;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const float* xr = x + (size_t)r * d;
        const float x2 = __double2float_rn(canon_dot_global(xr, xr, dq, lane));
        for (int c = 0; c < k; ++c) {
            const float* ec = cb + (size_t)c * d;
            const double s = canon_dot_global(xr, ec, dq, lane);
            const float e2c = __double2float_rn(canon_dot_global(ec, ec, dq, lane));
            if (lane == 0) dist[(size_t)r * k + c] = -canon_score(x2, s, e2c);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Dead-code re-seed (vq.py:181-195): only `embed` rows whose EMA cluster size fell below the
// threshold are replaced by the sampled batch rows.
__global__ void __launch_bounds__(256) reseed_kernel(const float* __restrict__ x, const int64_t* __restrict__ rows,
                                                      const float* __restrict__ cluster_size, float threshold,
                                                      float* __restrict__ embed, int64_t n, int k, int d) {
    const int lane = threadIdx.x & 31;
    const int dq = d >> 2;
    for (int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < k; c += (gridDim.x * blockDim.x) >> 5) {
        if (!(__ldg(cluster_size + c) < threshold)) continue;
        int64_t r = __ldg(rows + c);
        r = r < 0 ? 0 : (r >= n ? n - 1 : r);
        const float4* xr = reinterpret_cast<const float4*>(x + (size_t)r * d);
        float4* er = reinterpret_cast<float4*>(embed + (size_t)c * d);
        for (int q = lane; q < dq; q += 32) er[q] = __ldg(xr + q);
    }
}

}  // namespace tvq
