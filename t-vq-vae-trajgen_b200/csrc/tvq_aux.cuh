// tvq_aux.cuh — the small kernels around the fused forward: per-call preparation, EMA codebook
// update, backward, de-tokenising gather, dense distance matrix and dead-code re-seed.
// All of them are HBM- or latency-bound elementwise / gather work (no tensor cores).
#pragma once
#include <cuda_bf16.h>

#include "tvq_common.cuh"

namespace tvq {

// ------------------------------------------------------------------------------------------
// Per-call preparation: canonical |e_k|^2 (one warp per code), zero the statistics buffer and
// the loss / diagnostic fields of the workspace header; optionally the operands the streamed tcgen05 path feeds to TMA:
//   cbh  [k, dp] bf16 row-major, zero padded to dp columns: the NEGATED codebook, -e
//   e2h  [roundup(k, 256), 16] bf16: |e_k|^2 / 2 split into three bf16 pieces (hi, mid, lo: 24 bits, i.e. the fp32 value
//        exactly) followed by zeros — one extra K step of the distance GEMM against constant rows (1, 1, 1, 0...), so the
//        accumulator comes out as the half-score |e|^2 / 2 - x.e; rows >= k hold a huge value and can never be nominated
__global__ void __launch_bounds__(256) prep_kernel(const float* __restrict__ cb, int k, int d, float* __restrict__ e2,
                                                   WsHeader* hdr, float* __restrict__ stats, int64_t stats_len,
                                                   __nv_bfloat16* __restrict__ cbh, int dp, __nv_bfloat16* __restrict__ e2h) {
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int kpad = (k + 255) & ~255;
    for (int c = gwarp; c < kpad; c += nwarps) {
        float e2c = 1e30f;                                               // pad: never the minimum
        if (c < k) {
            const float* er = cb + (size_t)c * d;
            e2c = __double2float_rn(canon_dot_global(er, er, d >> 2, lane));
        }
        if (lane == 0) e2[c] = e2c;
        // |e_c - bf16(e_c)|^2, rounded UP to a bf16: the streamed kernel's error bound uses the ACTUAL rounding error of
        // its bf16 operands (max over codes), not the worst case 2^-9 |e|.  It travels in column 3 of the |e|^2 / 2 slab,
        // where the constant A rows of that K step hold a zero (0 x finite = 0).
        float de2c = 0.f;
        if (e2h != nullptr && c < k) {
            const float4* er4 = reinterpret_cast<const float4*>(cb + (size_t)c * d);
            for (int c4 = lane; c4 < (d >> 2); c4 += 32) {
                const float4 v = __ldg(er4 + c4);
                const float tx = v.x - __bfloat162float(__float2bfloat16_rn(v.x)), ty = v.y - __bfloat162float(__float2bfloat16_rn(v.y));
                const float tz = v.z - __bfloat162float(__float2bfloat16_rn(v.z)), tw = v.w - __bfloat162float(__float2bfloat16_rn(v.w));
                de2c = fmaf(tx, tx, fmaf(ty, ty, fmaf(tz, tz, fmaf(tw, tw, de2c))));
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) de2c += __shfl_xor_sync(0xffffffffu, de2c, off);
            de2c = fminf(de2c * 1.0001f, 3e38f);
            if (!(de2c >= 0.f)) de2c = 3e38f;                            // NaN code word: the latent bound becomes huge, never NaN x 0
        }
        if (e2h != nullptr && lane < 16) {
            const float v = 0.5f * e2c;                                  // exact
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            const float r1 = v - __bfloat162float(h);                    // exact (Sterbenz-like: h is v rounded to 8 bits)
            const __nv_bfloat16 m = __float2bfloat16_rn(r1);
            const __nv_bfloat16 l = __float2bfloat16_rn(r1 - __bfloat162float(m));
            e2h[(size_t)c * 16 + lane] = lane == 0 ? h : lane == 1 ? m : lane == 2 ? l : lane == 3 ? __float2bfloat16_ru(de2c) : __float2bfloat16_rn(0.f);
        }
    }
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
            const __nv_bfloat162 lo = __floats2bfloat162_rn(-v.x, -v.y), hi = __floats2bfloat162_rn(-v.z, -v.w);
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
// Data-parallel EMA update in ONE kernel: a one-shot all-reduce of the packed statistics over NVLink
// peer memory, fused with the EMA update (the reference's two dormant all_reduce hooks, vq.py:229 and
// :234, followed by :231 and :236-242).  Every rank owns an exchange buffer that all peers have mapped:
//   header (64 B, local only: u32 launch counter, u32 error word = last step a peer timed out in) | flags [2 parities][world] u32 | slots [2][world][len4] fp32
// Step e (parity e & 1): push my statistics into slot [parity][rank] of EVERY rank's buffer (posted
// remote stores), fence, publish flag = e on every rank, wait until all `world` local flags show e, then
// add the slots in rank order — the same order on every rank, so the replicas stay bit-identical — and
// apply the update.  Two parities suffice: a rank can only run one step ahead of the slowest peer
// (its next step needs that peer's flag of the current one).
// One CTA (statistics of <= 64 Ki floats); the launch counter lives in device memory so the kernel can be
// replayed from a CUDA graph.
struct EmaDpParams {
    EmaParams e;                 // e.stats = this rank's local statistics (read), buffers updated in place
    void* const* peers;          // device array [world]: every rank's exchange buffer
    int rank, world;
    int64_t len4;                // statistics length in float4 (padded)
    unsigned long long timeout_ns;   // give up waiting for a peer after this long (0 = never); see wait_flag_sys
    int finalize;                // 1: the statistics of this step were already published by the forward kernel
                                 // (tvq_fwd_simt.cuh::dp_publish_stats): no push, the step counter is read, not advanced
    float* consume;              // optional: e.stats is the workspace scratch of the fused step: zero it once it is pushed
};

__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(1024) ema_dp_kernel(const EmaDpParams p) {
    __shared__ double red[32];
    __shared__ float s_n;
    __shared__ unsigned s_epoch;
    const int tid = threadIdx.x;
    unsigned char* mine = reinterpret_cast<unsigned char*>(p.peers[p.rank]);
    if (tid == 0) {
        unsigned* counter = reinterpret_cast<unsigned*>(mine);
        s_epoch = *reinterpret_cast<volatile unsigned*>(counter) + (p.finalize ? 0u : 1u);
        if (!p.finalize) *counter = s_epoch;
    }
    __syncthreads();
    const unsigned epoch = s_epoch;
    const int par = (int)(epoch & 1u);
    const size_t flags_off = 64, slots_off = 64 + (((size_t)2 * p.world * 4 + 63) & ~(size_t)63);
    // ---- push
    if (!p.finalize) {
        const float4* src = reinterpret_cast<const float4*>(p.e.stats);
        for (int r = 0; r < p.world; ++r) {
            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(p.peers[r]) + slots_off) +
                          ((size_t)par * p.world + p.rank) * p.len4;
            for (int64_t f = tid; f < p.len4; f += blockDim.x) dst[f] = src[f];
        }
        __threadfence_system();
        if (p.consume) {         // (every thread zeroes exactly the cells it has just read)
            float4* z = reinterpret_cast<float4*>(p.consume);
            for (int64_t f = tid; f < p.len4; f += blockDim.x) z[f] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    __syncthreads();
    if (tid < p.world) {
        unsigned* flag = reinterpret_cast<unsigned*>(reinterpret_cast<unsigned char*>(p.peers[tid]) + flags_off) + par * p.world + p.rank;
        if (!p.finalize) st_release_sys_u32(flag, epoch);
        // ---- wait for every rank's contribution to MY buffer (bounded: a lost peer becomes a reported error, not a hang)
        const unsigned* lf = reinterpret_cast<const unsigned*>(mine + flags_off) + par * p.world + tid;
        wait_flag_sys(lf, epoch, p.timeout_ns, reinterpret_cast<unsigned*>(mine) + 1);
    }
    __syncthreads();
    // ---- reduce (rank order) + EMA update
    const float4* slots = reinterpret_cast<const float4*>(mine + slots_off) + (size_t)par * p.world * p.len4;
    const int kp = (p.e.k + 3) & ~3;
    const int dq = p.e.d >> 2;
    double part = 0.0;
    for (int c = tid; c < p.e.k; c += blockDim.x) {
        float cnt = 0.f;
        for (int r = 0; r < p.world; ++r) cnt += __ldcg(reinterpret_cast<const float*>(slots + (size_t)r * p.len4) + c);
        part += (double)ema_mix(p.e.cluster_size[c], cnt, p.e.decay, p.e.one_minus_decay);
    }
    double tot = block_sum(part, red);
    if (tid == 0) s_n = __double2float_rn(tot);
    __syncthreads();
    const float n = s_n;
    const float denom = __fadd_rn(n, p.e.k_eps);
    float4* avg4 = reinterpret_cast<float4*>(p.e.embed_avg);
    float4* emb4 = reinterpret_cast<float4*>(p.e.embed);
    float4* prev4 = reinterpret_cast<float4*>(p.e.embed_prev);
    for (int64_t f = tid; f < (int64_t)p.e.k * dq; f += blockDim.x) {
        const int c = (int)(f / dq);
        float cnt = 0.f;
        float4 sv = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < p.world; ++r) {
            const float4* sl = slots + (size_t)r * p.len4;
            cnt += __ldcg(reinterpret_cast<const float*>(sl) + c);
            const float4 v = __ldcg(sl + (kp >> 2) + f);
            sv.x += v.x; sv.y += v.y; sv.z += v.z; sv.w += v.w;
        }
        const float cs = ema_mix(p.e.cluster_size[c], cnt, p.e.decay, p.e.one_minus_decay);
        const float sm = __fmul_rn(__fdiv_rn(__fadd_rn(cs, p.e.eps), denom), n);
        float4 a = avg4[f];
        a.x = ema_mix(a.x, sv.x, p.e.decay, p.e.one_minus_decay);
        a.y = ema_mix(a.y, sv.y, p.e.decay, p.e.one_minus_decay);
        a.z = ema_mix(a.z, sv.z, p.e.decay, p.e.one_minus_decay);
        a.w = ema_mix(a.w, sv.w, p.e.decay, p.e.one_minus_decay);
        avg4[f] = a;
        if (prev4) prev4[f] = emb4[f];
        emb4[f] = make_float4(__fdiv_rn(a.x, sm), __fdiv_rn(a.y, sm), __fdiv_rn(a.z, sm), __fdiv_rn(a.w, sm));
    }
    __syncthreads();                                      // every thread has read the old cluster sizes
    for (int c = tid; c < p.e.k; c += blockDim.x) {
        float cnt = 0.f;
        for (int r = 0; r < p.world; ++r) cnt += __ldcg(reinterpret_cast<const float*>(slots + (size_t)r * p.len4) + c);
        p.e.cluster_size[c] = ema_mix(p.e.cluster_size[c], cnt, p.e.decay, p.e.one_minus_decay);
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
                                                           int64_t ntok, int k, int d, float* __restrict__ out,
                                                           unsigned* __restrict__ bad) {
    // An id outside [0, k) (e.g. a leaked mask token, id == k) is an error F.embedding raises on: its row is written as
    // NaN and counted in *bad (if given) — never silently decoded as a neighbouring code.
    const int lane = threadIdx.x & 31;
    const int dq = d >> 2;
    const float nan = __int_as_float(0x7fc00000);
    for (int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < ntok; t += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const int64_t code = __ldg(tok + t);
        if (code < 0 || code >= k) {
            if (lane == 0 && bad) atomicAdd(bad, 1u);
            for (int c = lane; c < dq; c += 32) st_stream_v4(out + (size_t)t * d + 4 * c, make_float4(nan, nan, nan, nan));
            continue;
        }
        const float4* er = reinterpret_cast<const float4*>(cb + (size_t)code * d);
        for (int c = lane; c < dq; c += 32) st_stream_v4(out + (size_t)t * d + 4 * c, __ldg(er + c));
    }
}
// layout 1: out[b, :, t] = cb[tok[b, t]]  (decoder layout "b c (h w)"): 32 tokens x 32 channels
// are transposed through shared memory so that both the code-word reads and the output writes
// are coalesced.
__global__ void __launch_bounds__(256) gather_transposed_kernel(const int64_t* __restrict__ tok, const float* __restrict__ cb,
                                                                 int64_t b, int64_t t, int k, int d, float* __restrict__ out,
                                                                 unsigned* __restrict__ bad) {
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
            const bool oob = code < 0 || code >= k;          // -> NaN column + error count (see gather_rows_kernel)
            if (oob && bad && d0 == 0) atomicAdd(bad, 1u);
            codes[threadIdx.x] = oob ? -1 : (int)code;
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8)   // r: token within tile, tx: channel
            tile[r][tx] = (d0 + tx < d) ? (codes[r] < 0 ? __int_as_float(0x7fc00000) : __ldg(cb + (size_t)codes[r] * d + d0 + tx)) : 0.f;
        __syncthreads();
        for (int r = ty; r < 32; r += 8)   // r: channel within tile, tx: token
            if (d0 + r < d && t0 + tx < t) out[(bi * d + d0 + r) * t + t0 + tx] = tile[tx][r];
    }
}

// ------------------------------------------------------------------------------------------
// Backward of the train forward for a caller that works in 'b c (h w)' (utils/train_utils.py:346-349): the gradient
// w.r.t. z_q arrives channels-first, the gradient w.r.t. z must leave channels-first, while x (the latents the forward
// saw) and idx are row-major.  One kernel instead of [transpose g, backward_kernel, transpose g_x]: 32 x 32 tiles are
// turned through shared memory, so every global access is a coalesced 128-byte row.
//   g_z[b, c, t] = g_zq[b, c, t] + coef * (x[b t, c] - q_st[b t, c]),  q_st = x + (e[idx] - x)
__global__ void __launch_bounds__(256) backward_cf_kernel(const float* __restrict__ g_zq, const float* __restrict__ g_commit,
                                                           const float* __restrict__ g_weighted, const float* __restrict__ x,
                                                           const int64_t* __restrict__ idx, const float* __restrict__ cb, int64_t b,
                                                           int hw, int d, float weight, float scale, float* __restrict__ g_z) {
    __shared__ float tg[32][33];
    __shared__ float to[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float coef = fmaf(weight, g_weighted ? __ldg(g_weighted) : 0.f, g_commit ? __ldg(g_commit) : 0.f) * scale;
    const int64_t tt = (hw + 31) / 32, ct = (d + 31) / 32;
    const int64_t total = b * tt * ct;
    for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
        const int64_t bi = w / (tt * ct);
        const int64_t rem = w - bi * tt * ct;
        const int t0 = (int)(rem / ct) * 32, c0 = (int)(rem % ct) * 32;
        __syncthreads();
        if (g_zq) {
#pragma unroll
            for (int i = ty; i < 32; i += 8)       // i: channel, tx: position — coalesced along (h w)
                tg[i][tx] = (c0 + i < d && t0 + tx < hw) ? ld_stream_v1(g_zq + (bi * d + c0 + i) * hw + t0 + tx) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int i = ty; i < 32; i += 8) {         // i: position, tx: channel — coalesced along c
            float r = 0.f;
            if (t0 + i < hw && c0 + tx < d) {
                const int64_t row = bi * hw + t0 + i;
                const float xv = ld_stream_v1(x + row * d + c0 + tx);
                const float ev = __ldg(cb + (size_t)__ldg(idx + row) * d + c0 + tx);
                const float qst = __fadd_rn(xv, __fsub_rn(ev, xv));
                r = fmaf(coef, __fsub_rn(xv, qst), g_zq ? tg[tx][i] : 0.f);
            }
            to[tx][i] = r;
        }
        __syncthreads();
#pragma unroll
        for (int i = ty; i < 32; i += 8)           // i: channel, tx: position
            if (c0 + i < d && t0 + tx < hw) g_z[(bi * d + c0 + i) * hw + t0 + tx] = to[i][tx];
    }
}

// The same, one CTA per batch element (used when its x slab and the codebook fit in shared memory: always at the
// shipped shapes).  Per batch element g_zq[b], x[b] and g_z[b] are CONTIGUOUS blocks of hw*d floats: x[b] is read with
// coalesced 16-byte loads into a padded shared-memory slab (row stride d + 1: a column read is conflict-free), the
// codebook likewise, and the output is produced in g's own (c, t) order — two phases instead of three per 32 x 32 tile.
__global__ void __launch_bounds__(256) backward_cf_slab_kernel(const float* __restrict__ g_zq, const float* __restrict__ g_commit,
                                                                const float* __restrict__ g_weighted, const float* __restrict__ x,
                                                                const int64_t* __restrict__ idx, const float* __restrict__ cb,
                                                                int64_t b, int hw, int k, int d, float weight, float scale,
                                                                float* __restrict__ g_z) {
    extern __shared__ float bsm[];
    const int ds = d + 1;
    float* xs = bsm;                                     // [hw][d + 1]
    float* es = xs + hw * ds;                            // [k][d + 1]
    int* cs = reinterpret_cast<int*>(es + k * ds);       // [hw] codes
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const float coef = fmaf(weight, g_weighted ? __ldg(g_weighted) : 0.f, g_commit ? __ldg(g_commit) : 0.f) * scale;
    const int dq = d >> 2;
    for (int f = tid; f < k * dq; f += blockDim.x) {     // codebook -> padded shared memory (once per CTA)
        const int row = f / dq, c4 = f - row * dq;
        const float4 v = __ldg(reinterpret_cast<const float4*>(cb + (size_t)row * d) + c4);
        float* dst = es + row * ds + 4 * c4;
        dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    }
    for (int64_t bi = blockIdx.x; bi < b; bi += gridDim.x) {
        __syncthreads();
        const float4* xb = reinterpret_cast<const float4*>(x + bi * (int64_t)hw * d);
#pragma unroll 4
        for (int f = tid; f < hw * dq; f += blockDim.x) {
            const int row = f / dq, c4 = f - row * dq;
            const float4 v = ld_stream_v4(reinterpret_cast<const float*>(xb + f));
            float* dst = xs + row * ds + 4 * c4;
            dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
        }
        for (int t = tid; t < hw; t += blockDim.x) {
            int64_t code = __ldg(idx + bi * hw + t);
            cs[t] = (int)(code < 0 ? 0 : (code >= k ? k - 1 : code));
        }
        __syncthreads();
        const float* gb = g_zq ? g_zq + bi * (int64_t)hw * d : nullptr;
        float* ob = g_z + bi * (int64_t)hw * d;
        // four channels per warp pass, positions across lanes: the (independent) g loads of a pass are all in flight
        // before the first one is used
        for (int c0 = 4 * warp; c0 < d; c0 += 4 * nwarps) {
            for (int t0 = 0; t0 < hw; t0 += 64) {
                float g[4][2];
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int t = t0 + lane + 32 * u;
                        g[j][u] = (gb && t < hw && c0 + j < d) ? ld_stream_v1(gb + (c0 + j) * hw + t) : 0.f;
                    }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int t = t0 + lane + 32 * u;
                    if (t >= hw) continue;
                    const int code = cs[t];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (c0 + j >= d) continue;
                        const float xv = xs[t * ds + c0 + j];
                        const float ev = es[code * ds + c0 + j];
                        const float qst = __fadd_rn(xv, __fsub_rn(ev, xv));
                        ob[(c0 + j) * hw + t] = fmaf(coef, __fsub_rn(xv, qst), g[j][u]);
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Backward with EVERYTHING channels-first (the forward read z in place, tvq_forward_cf / tvq_train_step_cf):
//   g_z[b, c, t] = g_zq[b, c, t] + coef * (z[b, c, t] - q_st[b, c, t]),  q_st = z + (e[idx[b, t]][c] - z)
// Per batch element z[b], g_zq[b] and g_z[b] are contiguous blocks of d * hw floats, so the kernel is a flat 16-byte
// streaming pass; the code words come from a padded shared-memory copy of the codebook (row stride d + 1: the lanes of
// a warp hold consecutive positions t, i.e. different codes, of mostly one channel).  12 d bytes per latent.
template <bool VEC, typename CodeT>
__global__ void __launch_bounds__(256) backward_cfx_kernel(const float* __restrict__ g_zq, const float* __restrict__ g_commit,
                                                            const float* __restrict__ g_weighted, const float* __restrict__ z,
                                                            const int64_t* __restrict__ idx, const float* __restrict__ cb, int64_t b,
                                                            int hw, int k, int d, float weight, float scale,
                                                            float* __restrict__ g_z) {
    extern __shared__ float bsm[];
    const int ds = d + 1;
    float* es = bsm;                                     // [k][d + 1]: distinct codes of one channel -> distinct banks (k <= 32)
    CodeT* cs = reinterpret_cast<CodeT*>(es + k * ds);   // [hw] codes of the current batch element (bytes for k <= 256:
                                                         //  the 4 consecutive positions of a lane share a word, lanes differ)
    const int tid = threadIdx.x;
    const float coef = fmaf(weight, g_weighted ? __ldg(g_weighted) : 0.f, g_commit ? __ldg(g_commit) : 0.f) * scale;
    for (int f = tid; f < k * d; f += blockDim.x) es[(f / d) * ds + (f % d)] = __ldg(cb + f);
    const int slab = hw * d;
    // element 4 * (tid + i * blockDim) = channel c, position t; one step of the loop advances it by 4 * blockDim elements
    const int step_c = (4 * (int)blockDim.x) / hw, step_t = 4 * (int)blockDim.x - step_c * hw;
    for (int64_t bi = blockIdx.x; bi < b; bi += gridDim.x) {
        __syncthreads();                                 // the previous element's codes are no longer read
        for (int t = tid; t < hw; t += blockDim.x) {
            const int64_t code = __ldg(idx + bi * hw + t);
            cs[t] = (CodeT)(code < 0 ? 0 : (code >= k ? k - 1 : code));
        }
        __syncthreads();
        const float* zb = z + bi * (int64_t)slab;
        const float* gb = g_zq ? g_zq + bi * (int64_t)slab : nullptr;
        float* ob = g_z + bi * (int64_t)slab;
        if (VEC) {
            int c0 = (4 * tid) / hw, t0 = 4 * tid - c0 * hw;
            for (int f = tid; f < (slab >> 2); f += blockDim.x) {
                const float4 xv = ld_stream_v4(zb + 4 * f);
                const float4 gv = gb ? ld_stream_v4(gb + 4 * f) : make_float4(0.f, 0.f, 0.f, 0.f);
                int c = c0, t = t0;
                const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, gs[4] = {gv.x, gv.y, gv.z, gv.w};
                float o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float ev = es[(int)cs[t] * ds + c];
                    const float qst = __fadd_rn(xs[j], __fsub_rn(ev, xs[j]));
                    o[j] = fmaf(coef, __fsub_rn(xs[j], qst), gs[j]);
                    if (++t == hw) { t = 0; ++c; }
                }
                st_stream_v4(ob + 4 * f, make_float4(o[0], o[1], o[2], o[3]));
                c0 += step_c; t0 += step_t;
                if (t0 >= hw) { t0 -= hw; ++c0; }
            }
        } else {
            for (int f = tid; f < slab; f += blockDim.x) {
                const int c = f / hw, t = f - c * hw;
                const float xv = ld_stream_v1(zb + f);
                const float ev = es[(int)cs[t] * ds + c];
                const float qst = __fadd_rn(xv, __fsub_rn(ev, xv));
                ob[f] = fmaf(coef, __fsub_rn(xv, qst), gb ? ld_stream_v1(gb + f) : 0.f);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Batched 2-D transpose in[b][r][s] -> out[b][s][r]: the layout change of quantize() (utils/train_utils.py:346-349:
// 'b c h w -> b (h w) c' before the VQ and back after it) as a shared-memory-tiled copy — both the reads (along s)
// and the writes (along r) are coalesced 128-byte rows, where the generic strided copy torch runs for
// rearrange(...).contiguous() reaches about a third of that.  8d bytes per latent per direction.
__global__ void __launch_bounds__(256) batched_transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t b,
                                                                 int r, int s) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 8 warps
    const int64_t rt = (r + 31) / 32, st = (s + 31) / 32;
    const int64_t total = b * rt * st;
    for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
        const int64_t bi = w / (rt * st);
        const int64_t rem = w - bi * rt * st;
        const int r0 = (int)(rem / st) * 32, s0 = (int)(rem % st) * 32;
        const float* src = in + bi * (int64_t)r * s;
        float* dst = out + bi * (int64_t)r * s;
        __syncthreads();
#pragma unroll
        for (int i = ty; i < 32; i += 8)
            if (r0 + i < r && s0 + tx < s) tile[i][tx] = ld_stream_v1(src + (int64_t)(r0 + i) * s + s0 + tx);
        __syncthreads();
#pragma unroll
        for (int i = ty; i < 32; i += 8)
            if (s0 + i < s && r0 + tx < r) dst[(int64_t)(s0 + i) * r + r0 + tx] = tile[tx][i];
    }
}

// The same transpose with one batch element per CTA pass: in[b] (r x s floats, contiguous) is read front to back
// into a padded shared-memory slab (row pitch odd: both the row-wise fill and the column-wise drain are conflict-free)
// and out[b] (s x r, contiguous) is written front to back — every global access is a full 128-byte line whatever r and
// s are (the 32 x 32 tiling leaves 44 % of a tile empty at s = 18).  Used for r or s below 32 when the slab fits.
__global__ void __launch_bounds__(256) slab_transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t b,
                                                              int r, int s) {
    extern __shared__ float slab[];
    const int P = s | 1;                                  // row pitch (odd)
    const int n = r * s, tid = threadIdx.x, nt = blockDim.x;
    // element tid + i * nt of the input is (c, t) = divmod(., s); of the output (t, c) = divmod(., r): incremental
    const int ic = nt / s, it = nt - ic * s, oc = nt / r, ot = nt - oc * r;
    for (int64_t bi = blockIdx.x; bi < b; bi += gridDim.x) {
        const float* src = in + bi * (int64_t)n;
        float* dst = out + bi * (int64_t)n;
        __syncthreads();                                  // the previous element has been drained
        int c = tid / s, t = tid - c * s;
        for (int e = tid; e < n; e += nt) {
            slab[c * P + t] = ld_stream_v1(src + e);
            c += ic; t += it;
            if (t >= s) { t -= s; ++c; }
        }
        __syncthreads();
        int t2 = tid / r, c2 = tid - t2 * r;
        for (int e = tid; e < n; e += nt) {
            dst[e] = slab[c2 * P + t2];
            t2 += oc; c2 += ot;
            if (c2 >= r) { c2 -= r; ++t2; }
        }
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
