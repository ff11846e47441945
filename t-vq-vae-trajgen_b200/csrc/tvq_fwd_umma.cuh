// tvq_fwd_umma.cuh — fused VQ forward with tcgen05 scoring, codebook resident in shared memory
// (train: k <= 32, eval: k <= 64; d <= 128).  This is the path BASELINE configs[0,1,3,4] run
// (k = 32, d = 128).
//
// One persistent CTA per SM, 10 warps, warp-specialised:
//   warp 0      TMA producer: x tiles of 64 latents x d fp32 -> ring of shared-memory stages
//               (cp.async.bulk.tensor, SWIZZLE_128B, zero-fill past n / past d)
//   warp 1      MMA issuer: per tile d/8 tcgen05.mma.kind::tf32 (M=64, N=KP) straight from the fp32
//               tiles (no conversion pass) into one of 4 TMEM accumulator slots; tcgen05.commit
//   warps 2-9   eight epilogue warps in two groups of four (one warp per TMEM lane quadrant); the
//               groups take alternate tiles.  A warp owns the 16 rows of its quadrant from the
//               TMEM read to the last store, so epilogue warps never synchronise with each other:
//               the only hand-offs are mbarriers (tile landed / scores ready / slot free / stage free).
// Per tile an epilogue warp
//   1. reads its 16 x KP approximate dot products from TMEM (tcgen05.ld), forms scores
//      e2 - 2 x.e and keeps the four best (score, code) pairs of every row with a branch-free
//      min/max network on packed keys;
//   2. per row (all 32 lanes on one row, four rows in flight): |x|^2 -> the RIGOROUS tf32 error
//      bound -> rows with more than one code inside the bound are decided by
//      the canonical fp64 re-score (tvq_common.cuh, mirrored by oracle/vq_canon.c), so the indices
//      are exactly those of the SIMT path and of the C oracle whatever the tensor-core rounding did;
//   3. gathers the code word from the shared-memory codebook, forms the straight-through output
//      and the commitment-loss partial, streams q_st out (st.global.cs), writes idx;
//   4. adds the row into the warp's PRIVATE per-code accumulators, which live in tensor memory
//      (tcgen05.ld / tcgen05.st read-modify-write of 4 columns x 32 lanes): no shared-memory
//      accumulators, no sort, no floating-point atomics until the CTA's single flush;
//   5. releases the stage to the producer.
// Algorithmic HBM traffic: read x once, write q once, write idx: 8d + 8 bytes per latent.
#pragma once
#include <cuda.h>

#include "tvq_common.cuh"
#include "tvq_fwd_simt.cuh"
#include "tvq_sm100.cuh"

namespace tvq {

#ifdef TVQ_PROFILE_PHASES
// Experiment build only (tools/profile_phases.py): per-phase clock64 totals of CTA 0's epilogue warps.
__device__ unsigned long long g_phase_clk[2][16];
__device__ unsigned long long g_phase_all[2][12];  // CTA 0, epilogue groups 0/1, quadrant 0: clock totals of phases 0..9
__device__ unsigned long long g_tile_clk[8][4];   // CTA 0, warp 2: per tile (first 8): clocks at tile landed / scores ready / scan done / apply done
__device__ unsigned long long g_gt[4];   // [0] min CTA start (globaltimer ns), [1] max CTA end, [2] max CTA clocks, [3] max main-loop-end clocks
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define TVQ_PH(i) do { long long _t = clock64(); ph_acc[i] += _t - ph_t; ph_t = _t; } while (0)
#else
#define TVQ_PH(i) do { } while (0)
#endif

#ifndef TVQ_UGROUPS
#define TVQ_UGROUPS 3
#endif
constexpr int kUM = 64;                 // latents per UMMA tile (M)
constexpr int kUGroups = TVQ_UGROUPS;   // epilogue groups of 4 warps (one warp per TMEM lane quadrant)
constexpr int kUThreads = 64 + 128 * kUGroups + 32;   // producer warp + MMA warp + epilogue warps + second loader warp (XCF)
constexpr int kUWarpL2 = 2 + 4 * kUGroups;       // the second loader warp of the channels-first variant (idle otherwise)
constexpr int kXcfDepth = 3;                     // channels-first loader: tiles in flight behind the one being issued
constexpr int kUSlots = 4;              // TMEM score slots
constexpr int kUMaxStages = 8;
constexpr int kUBatch = 4;              // rows a warp keeps in flight in the apply phase

#ifdef TVQ_USPLIT
#define TVQ_USPLIT_PLAN TVQ_USPLIT
#else
#define TVQ_USPLIT_PLAN 1
#endif
struct UmmaPlan {
    int stages, stage_bytes;
    int x, cb, cbh, cbl, e2s, hist, keys, red, misc, tiles, bars, tmem, total;
};
__host__ __device__ inline UmmaPlan make_umma_plan(int dp, int kp, int stages) {
    UmmaPlan u;
    u.stages = stages;
    u.stage_bytes = kUM * dp * 4;
    int o = 0;
    u.x = o;    o += stages * u.stage_bytes;
    u.cb = o;   o += kp * dp * 4;           // exact code words (gather, re-score)
    // k <= 32 (every training shape): the tensor core reads the code words as hi + lo (kUSplit)
    u.cbh = o;  o += (TVQ_USPLIT_PLAN && kp <= 32) ? kp * dp * 4 : 0;   // B operand, first MMA: the code words cut to tf32 (19 leading bits)
    u.cbl = o;  o += (TVQ_USPLIT_PLAN && kp <= 32) ? kp * dp * 4 : 0;   // B operand, second MMA: what the cut removed (exact remainder)
    u.e2s = o;  o += kp * 4;
    u.hist = o; o += kp * 4;
    o = (o + 15) & ~15;
    u.keys = o; o += 4 * kUGroups * 16 * 16;   // per epilogue warp: 16 rows x (4 best keys)
    u.red = o;  o += 16 * 8;
    u.misc = o; o += 16 * 4;
    u.tiles = o; o += kUMaxStages * 4;          // tile id held by each stage (dynamic scheduler; -1 = no more tiles)
    u.bars = o; o += (2 * kUMaxStages + 2 * kUSlots) * 8;
    u.tmem = o; o += 16;
    u.total = o;
    return u;
}

// Sorted insert of `key` into (t0 <= t1 <= t2 <= t3): branch-free min/max network.
__device__ __forceinline__ void top4_insert(float key, float& t0, float& t1, float& t2, float& t3) {
    float c = fmaxf(t0, key);
    t0 = fminf(t0, key);
    float c2 = fmaxf(t1, c);
    t1 = fminf(t1, c);
    float c3 = fmaxf(t2, c2);
    t2 = fminf(t2, c2);
    t3 = fminf(t3, c3);
}

// Rows whose tf32 scores leave more than one code inside the error bound go through a cascade:
// level 2 re-scores the candidates with fp32 FMAs (error ~1e-6 relative, rigorous bound err32),
// and only if that still cannot separate them does level 3 apply the canonical fp64 rule.  All 32
// lanes work on the one row (lane l holds chunk l of x).  Returns the canonical argmin.
template <int KP>
__device__ __noinline__ int resolve_row(const float4 xv, const float bnd2, const unsigned c0, const unsigned c1,
                                           const unsigned c2, const int nc, const bool all_codes, const float* cbs,
                                           const float* e2s, const int k, const bool has_chunk, const int lane) {
    // return value: code | (1 << 16) if the decision needed the fp64 level
    const float thr32 = 1.3e-6f * bnd2;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!all_codes) {
        // three candidate slots (unused slots repeat the best code, which is harmless); everything is
        // kept in named registers — no dynamically indexed arrays, hence no local memory
        const unsigned a0 = c0, a1 = nc > 1 ? c1 : c0, a2 = nc > 2 ? c2 : c0;
        const float4 e0 = has_chunk ? *reinterpret_cast<const float4*>(cbs + tile_off<KP>((int)a0, lane)) : z4;
        const float4 e1 = has_chunk ? *reinterpret_cast<const float4*>(cbs + tile_off<KP>((int)a1, lane)) : z4;
        const float4 e2v = has_chunk ? *reinterpret_cast<const float4*>(cbs + tile_off<KP>((int)a2, lane)) : z4;
        float d0 = fmaf(xv.x, e0.x, fmaf(xv.y, e0.y, fmaf(xv.z, e0.z, xv.w * e0.w)));
        float d1 = fmaf(xv.x, e1.x, fmaf(xv.y, e1.y, fmaf(xv.z, e1.z, xv.w * e1.w)));
        float d2 = fmaf(xv.x, e2v.x, fmaf(xv.y, e2v.y, fmaf(xv.z, e2v.z, xv.w * e2v.w)));
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            d0 += __shfl_xor_sync(0xffffffffu, d0, off);
            d1 += __shfl_xor_sync(0xffffffffu, d1, off);
            d2 += __shfl_xor_sync(0xffffffffu, d2, off);
        }
        const float s0 = fmaf(-2.f, d0, e2s[a0]), s1 = fmaf(-2.f, d1, e2s[a1]), s2 = fmaf(-2.f, d2, e2s[a2]);
        float sb = s0;
        unsigned cb_ = a0;
        if (s1 < sb || (s1 == sb && a1 < cb_)) { sb = s1; cb_ = a1; }
        if (s2 < sb || (s2 == sb && a2 < cb_)) { sb = s2; cb_ = a2; }
        const float lim = sb + thr32;
        const bool amb = (a0 != cb_ && s0 <= lim) || (a1 != cb_ && s1 <= lim) || (a2 != cb_ && s2 <= lim);
        if (!amb) return (int)cb_;
        // level 3: canonical fp64 among the candidates
        const float x2 = __double2float_rn(butterfly_sum(dot4(0.0, xv, xv)));
        const float q0 = canon_score(x2, butterfly_sum(dot4(0.0, xv, e0)), e2s[a0]);
        const float q1 = canon_score(x2, butterfly_sum(dot4(0.0, xv, e1)), e2s[a1]);
        const float q2 = canon_score(x2, butterfly_sum(dot4(0.0, xv, e2v)), e2s[a2]);
        float best = q0;
        int arg = (int)a0;
        if (q1 < best || (q1 == best && (int)a1 < arg)) { best = q1; arg = (int)a1; }
        if (q2 < best || (q2 == best && (int)a2 < arg)) { best = q2; arg = (int)a2; }
        return arg | (1 << 16);
    }
    // candidate list overflowed (or non-finite scores): fp32 pass over every code, then fp64 if needed
    float m1 = __int_as_float(0x7f800000), m2 = m1;
    int i1 = 0;
    for (int c = 0; c < k; ++c) {
        const float4 ev = has_chunk ? *reinterpret_cast<const float4*>(cbs + tile_off<KP>(c, lane)) : z4;
        float dd = fmaf(xv.x, ev.x, fmaf(xv.y, ev.y, fmaf(xv.z, ev.z, xv.w * ev.w)));
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) dd += __shfl_xor_sync(0xffffffffu, dd, off);
        const float sv = fmaf(-2.f, dd, e2s[c]);
        if (sv < m1) { m2 = m1; m1 = sv; i1 = c; }
        else if (sv < m2) m2 = sv;
    }
    if (m2 - m1 > thr32) return i1;
    const float x2 = __double2float_rn(butterfly_sum(dot4(0.0, xv, xv)));
    float best = __int_as_float(0x7f800000);
    int arg = 0;
    for (int c = 0; c < k; ++c) {
        const float4 ev = has_chunk ? *reinterpret_cast<const float4*>(cbs + tile_off<KP>(c, lane)) : z4;
        const float dk = canon_score(x2, butterfly_sum(dot4(0.0, xv, ev)), e2s[c]);
        if (dk < best) { best = dk; arg = c; }
    }
    return arg | (1 << 16);
}

// Everything one epilogue warp needs to finish ONE row (slow, fully checked form).  Used for the
// rows of the last, partial tile only; the steady state is the batched code in the kernel body.
struct RowResult {
    float loss;
    unsigned counters;   // low 16 bits: re-scored (0/1); high 16 bits: needed fp64 (0/1)
};
template <int DP, int KP, bool TRAIN>
__device__ __noinline__ RowResult row_generic(float* q, int64_t* idx, const int d, const int k, const float* xt,
                                              const float* cbs, const float* e2s, int* hist, const float4 key, const int trow,
                                              const int64_t grow, const bool has_chunk, const int lane, const float emax,
                                              const float err_p, const float err_s, const uint32_t acc_base, const int q_hw) {
    unsigned counters = 0;
    using namespace sm100;
    const float BIG = 1e30f;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 xv = has_chunk ? *reinterpret_cast<const float4*>(xt + tile_off<kUM>(trow, lane)) : z4;
    float ss = fmaf(xv.x, xv.x, fmaf(xv.y, xv.y, fmaf(xv.z, xv.z, xv.w * xv.w)));
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    const float xn = sqrt_approx(ss) * 1.0001f;
    const float bnd = xn + emax;
    const float bnd2 = bnd * bnd;
    const float lim = fmaf(err_p, xn, fmaf(err_s, bnd2, key.x));
    int code = (int)(__float_as_uint(key.x) & 63u);
    if (!(key.y > lim) || !(key.x < BIG)) {
        const int nc = 1 + (key.y <= lim) + (key.z <= lim) + (key.w <= lim);
        const bool finite = key.x < BIG;
        const int r = resolve_row<KP>(xv, bnd2, __float_as_uint(key.x) & 63u, __float_as_uint(key.y) & 63u,
                                      __float_as_uint(key.z) & 63u, nc, nc == 4 || !finite, cbs, e2s, k, has_chunk, lane);
        counters += 1u + ((unsigned)(r >> 16) << 16);
        code = r & 0xffff;
        code = code < k ? code : 0;
    }
    float loss = 0.f;
    if (lane == 0) { atomicAdd(hist + code, 1); idx[grow] = (int64_t)code; }
    if (q != nullptr || TRAIN) {
        const float4 ev = has_chunk ? *reinterpret_cast<const float4*>(cbs + tile_off<KP>(code, lane)) : z4;
        float4 o = ev;
        if (TRAIN) {
            o.x = __fadd_rn(xv.x, __fsub_rn(ev.x, xv.x));
            o.y = __fadd_rn(xv.y, __fsub_rn(ev.y, xv.y));
            o.z = __fadd_rn(xv.z, __fsub_rn(ev.z, xv.z));
            o.w = __fadd_rn(xv.w, __fsub_rn(ev.w, xv.w));
            const float dx = __fsub_rn(o.x, xv.x), dy = __fsub_rn(o.y, xv.y), dz = __fsub_rn(o.z, xv.z), dw = __fsub_rn(o.w, xv.w);
            loss = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, dw * dw)));
        }
        if (q != nullptr && has_chunk) {
            if (q_hw > 0) {      // channels-first output [n / hw, d, hw]
                const int64_t bi = grow / q_hw, hw = grow - bi * q_hw;
                float* qc = q + (bi * d + 4 * lane) * q_hw + hw;
                qc[0] = o.x; qc[q_hw] = o.y; qc[2 * (size_t)q_hw] = o.z; qc[3 * (size_t)q_hw] = o.w;
            } else {
                st_stream_v4(q + (size_t)grow * d + 4 * lane, o);
            }
        }
    }
    if (TRAIN) {
        __syncwarp();
        tmem_st_wait();
        float4 a = tmem_ld_x4(acc_base + 4 * code);
        tmem_ld_wait();
        a.x += xv.x; a.y += xv.y; a.z += xv.z; a.w += xv.w;
        tmem_st_x4(acc_base + 4 * code, a);
    }
    RowResult rr;
    rr.loss = loss;
    rr.counters = counters;
    return rr;
}

// ---------------------------------------------------------------------------------------------
// Channels-first x: x is the caller's 'b c (h w)' tensor [n / hw, d, hw] (utils/train_utils.py:346-347
// reads it through a rearrange copy; here the forward kernel reads it in place).  TMA cannot address that view (the
// stride between channels, 4 * hw bytes, is not a multiple of 16), so two loader warps fill the SWIZZLE_128B tile with
// 4-byte cp.async copies (consecutive latents of one channel are contiguous in global memory).  Completion:
// cp.async.wait_group -> fence.proxy.async (the MMA reads the tile through the async proxy) -> mbarrier arrive.
__device__ __forceinline__ void cp_async4_zfill(uint32_t dst, const float* src, unsigned src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait_n() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One loader warp's half of a tile: lane = latent (rows 32*lw .. 32*lw+31), one channel per instruction — one or two
// 128-byte lines of global memory per instruction, 4-way bank conflicts on the shared-memory side.  (Measured against
// a lane = 8 latents x 4 channels mapping, which is conflict-free in shared memory but touches 4-8 lines: 71 vs 86 us
// per 76 800-latent launch.)
template <int DP>
__device__ __forceinline__ void xcf_issue_tile(const FwdParams& p, const uint32_t stage_base, const int tile, const int lw,
                                               const int lane) {
    const unsigned hw = (unsigned)p.x_hw;
    const int row = 32 * lw + lane;
    const int64_t n = (int64_t)tile * kUM + row;
    const bool rv = n < p.n;
    const unsigned nn = rv ? (unsigned)n : 0u;
    const unsigned bi = nn / hw, pos = nn - bi * hw;
    const float* src = p.x + (size_t)bi * p.d * hw + pos;
    const uint32_t dst_row = stage_base + (uint32_t)row * 128u;
    const uint32_t x7 = (uint32_t)(row & 7) << 4;
#pragma unroll 8
    for (int c = 0; c < DP; ++c) {
        const uint32_t dst = dst_row + (uint32_t)(c >> 5) * (kUM * 128) + ((((uint32_t)(c >> 2) & 7u) << 4) ^ x7) + ((uint32_t)(c & 3) << 2);
        const bool ok = rv && c < p.d;
        cp_async4_zfill(dst, ok ? src + (size_t)c * hw : p.x, ok ? 4u : 0u);
    }
}

// Last-CTA epilogue of the resident-codebook kernel (k <= 64; EMA: k <= 32, single rank): the same arithmetic as
// finish_ticket / ema_kernel, arranged for latency — at BASELINE configs[1] sizes the whole launch is ~30 us, so
// the serial tail matters.  embed_avg and cluster_size (which no CTA of this launch writes before the last ticket)
// are loaded BEFORE the ticket; after it there is ONE batch of L2 loads (counts, sums, loss), warp-level
// reductions (one code per lane) instead of block-wide ones, and the stores.
template <int KP, bool TRAIN>
__device__ __forceinline__ void finish_resident(const FwdParams& p, const float* cbs, int* misc) {
    const int tid = threadIdx.x, lane = tid & 31;
    const bool ema = TRAIN && p.fuse_ema;       // (data-parallel: p.dp_world > 1 implies the fused step)
    constexpr int J = (KP * 32 + kUThreads - 1) / kUThreads;       // float4 cells of the [k, d] buffers per thread
    const int dq = p.d >> 2, cells = p.k * dq, kp = (p.k + 3) & ~3;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 avg[J];
    float cs_old = 0.f;
    if (ema) {
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int f = tid + j * kUThreads;
            avg[j] = f < cells ? __ldcg(reinterpret_cast<const float4*>(p.embed_avg) + f) : z4;
        }
        if (lane < p.k) cs_old = __ldcg(p.cluster_size + lane);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned t = atomicAdd(&p.hdr->ticket, 1u);
        misc[1] = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!misc[1]) return;
    __threadfence();
    // ---- one batch of loads
    const float cnt0 = lane < p.k ? __ldcg(p.stats + lane) : 0.f;
    const float cnt1 = (KP > 32 && lane + 32 < p.k) ? __ldcg(p.stats + lane + 32) : 0.f;
    float4* esum4 = reinterpret_cast<float4*>(p.stats + kp);
    float4 sv[J];
    const bool dp = p.dp_world > 1;            // data-parallel: the sums are loaded after the exchange below
    if (ema && !dp) {
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int f = tid + j * kUThreads;
            sv[j] = f < cells ? __ldcg(esum4 + f) : z4;
        }
    }
    double ls = 0.0;
    unsigned n_resc = 0, n_ex = 0;
    if (tid == 0) {
        ls = *reinterpret_cast<volatile double*>(&p.hdr->loss_sum);
        n_resc = *reinterpret_cast<volatile unsigned*>(&p.hdr->n_rescored);
        n_ex = *reinterpret_cast<volatile unsigned*>(&p.hdr->n_exact);
    }
    // ---- perplexity = exp(-sum p log(p + 1e-10)), p = counts / n  (vq.py:246-247); every warp computes it
    const float fn = (float)p.n;
    const float pr0 = __fdiv_rn(cnt0, fn), pr1 = __fdiv_rn(cnt1, fn);
    double part = lane < p.k ? (double)(pr0 * logf(pr0 + 1e-10f)) : 0.0;
    if (KP > 32 && lane + 32 < p.k) part += (double)(pr1 * logf(pr1 + 1e-10f));
    const double tot = butterfly_sum(part);
    if (tid == 0) {
        p.scalars[1] = expf(-(float)tot);
        const float commit = TRAIN ? (float)(ls / ((double)p.n * (double)p.d)) : 0.f;
        p.scalars[0] = commit;
        p.scalars[2] = __fmul_rn(commit, p.commitment_weight);
        p.scalars[3] = 0.f;
        reinterpret_cast<unsigned*>(p.scalars)[4] = n_resc;
        reinterpret_cast<unsigned*>(p.scalars)[5] = n_ex;
        p.scalars[6] = 0.f;
        p.scalars[7] = 0.f;
        if (p.commit_out) *p.commit_out = commit;
        if (p.weighted_out) *p.weighted_out = __fmul_rn(commit, p.commitment_weight);
        p.hdr->ticket = 0;
        p.hdr->next_tile = 0u;
        p.hdr->loss_sum = 0.0;                 // consumed: the header is clean for the next call
        p.hdr->n_rescored = 0u;
        p.hdr->n_exact = 0u;
    }
    if (!ema) return;
    float cnt = cnt0;
    if (dp && p.dp_defer) {
        // deferred exchange: publish this rank's statistics and leave — the wait, the rank-ordered sum and the EMA update
        // run in tvq_ema_finalize_dp.  What only this kernel knows is saved here: the pre-update codebook for the backward;
        // the statistics scratch is zero again for the next call.
        // (dp_defer == 2: not even the publication happens here — the statistics stay in the workspace scratch, and
        // tvq_ema_finalize_dp pushes them itself and zeroes the scratch; this kernel's tail is then the single-GPU one)
        const bool publish = p.dp_defer == 1;
        if (publish) dp_publish_stats(p, misc);
        float4* prev4 = reinterpret_cast<float4*>(p.embed_prev);
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int f = tid + j * kUThreads;
            if (f < cells) {
                const int c = f / dq;
                if (prev4) prev4[f] = *reinterpret_cast<const float4*>(cbs + tile_off<KP>(c, f - c * dq));
                if (publish) esum4[f] = z4;
            }
        }
        if (publish && tid < kp) p.stats[tid] = 0.f;
        return;
    }
    if (dp) {
        // perplexity and loss above are this rank's own (as in the reference: vq.py:246-249 sits after, and is unaffected
        // by, the all-reduce); the EMA update uses the statistics of ALL ranks, summed over NVLink peer memory
        dp_reduce_stats(p, misc);
        cnt = lane < p.k ? __ldcg(p.stats + lane) : 0.f;
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int f = tid + j * kUThreads;
            sv[j] = f < cells ? __ldcg(esum4 + f) : z4;
        }
    }
    // ---- EMA update (vq.py:231,236-242): every other CTA has finished reading the codebook and flushing
    const float cs_new = lane < p.k ? fmaf(cnt, p.one_minus_decay, __fmul_rn(cs_old, p.decay)) : 0.f;
    const float nsum = __double2float_rn(butterfly_sum((double)cs_new));
    const float denom = __fadd_rn(nsum, p.k_eps);
    float4* avg4 = reinterpret_cast<float4*>(p.embed_avg);
    float4* emb4 = reinterpret_cast<float4*>(p.embed);
    float4* prev4 = reinterpret_cast<float4*>(p.embed_prev);
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int f = tid + j * kUThreads;
        const int c = f < cells ? f / dq : 0;
        const float cs = __shfl_sync(0xffffffffu, cs_new, c);
        if (f < cells) {
            const float sm = __fmul_rn(__fdiv_rn(__fadd_rn(cs, p.eps), denom), nsum);
            float4 a = avg[j];
            a.x = fmaf(sv[j].x, p.one_minus_decay, __fmul_rn(a.x, p.decay));
            a.y = fmaf(sv[j].y, p.one_minus_decay, __fmul_rn(a.y, p.decay));
            a.z = fmaf(sv[j].z, p.one_minus_decay, __fmul_rn(a.z, p.decay));
            a.w = fmaf(sv[j].w, p.one_minus_decay, __fmul_rn(a.w, p.decay));
            avg4[f] = a;
            if (prev4) prev4[f] = *reinterpret_cast<const float4*>(cbs + tile_off<KP>(c, f - c * dq));   // pre-update code word
            emb4[f] = make_float4(__fdiv_rn(a.x, sm), __fdiv_rn(a.y, sm), __fdiv_rn(a.z, sm), __fdiv_rn(a.w, sm));
            esum4[f] = z4;                     // statistics consumed: the scratch is zero for the next call
        }
    }
    __syncthreads();                           // every warp has read the counts
    if (tid < kp) {
        if (tid < p.k) p.cluster_size[tid] = cs_new;
        p.stats[tid] = 0.f;
    }
}

template <int DP, int KP, bool TRAIN, bool FULLD, bool XCF = false>
__global__ void __launch_bounds__(kUThreads, 1) fwd_umma_kernel(const __grid_constant__ CUtensorMap tmap_x, const FwdParams p,
                                                               const int stages) {
    using namespace sm100;
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int DPC = DP / 4;                 // 16-byte chunks per padded row (<= 32: one per lane)
    constexpr int NSLAB = DP / 32;              // 128-byte K slabs per tile
    constexpr int SLAB_X = kUM * 128;           // bytes of one x slab
    constexpr int SLAB_CB = KP * 128;           // bytes of one codebook slab
    constexpr int ACC_COLS = TRAIN ? kUGroups * 4 * KP : 0;   // per-warp accumulators: 4 columns per code
    constexpr uint32_t TMEM_NEED = kUSlots * KP + ACC_COLS;
    constexpr uint32_t TMEM_COLS = TMEM_NEED <= 32 ? 32 : TMEM_NEED <= 64 ? 64 : TMEM_NEED <= 128 ? 128 : TMEM_NEED <= 256 ? 256 : 512;
    static_assert(DP == 64 || DP == 128, "resident-codebook path: d padded to 64 or 128");
    static_assert(KP == 16 || KP == 32 || KP == 64, "resident-codebook path: k padded to 16, 32 or 64");
    static_assert(TMEM_NEED <= 512, "tensor memory budget");
    static_assert(kUBatch == 4, "the batched apply code is written for 4 rows in flight");

    const UmmaPlan pl = make_umma_plan(DP, KP, stages);
    float* cbs = reinterpret_cast<float*>(smem + pl.cb);
    float* cbhs = reinterpret_cast<float*>(smem + pl.cbh);
    float* cbls = reinterpret_cast<float*>(smem + pl.cbl);
    float* e2s = reinterpret_cast<float*>(smem + pl.e2s);
    int* hist = reinterpret_cast<int*>(smem + pl.hist);
    double* red = reinterpret_cast<double*>(smem + pl.red);
    int* misc = reinterpret_cast<int*>(smem + pl.misc);
    volatile int* stage_tile = reinterpret_cast<volatile int*>(smem + pl.tiles);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + pl.bars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + pl.tmem);
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * kUMaxStages;
    const uint32_t bar_tfull = bar_empty + 8 * kUMaxStages, bar_tempty = bar_tfull + 8 * kUSlots;
#ifndef TVQ_USPLIT
#define TVQ_USPLIT 1
#endif
    constexpr bool SPLIT = TVQ_USPLIT && KP <= 32;   // eval with 33..64 codes: one MMA on the exact code words (three 32 KB copies
                                                // would leave the tile ring 4 stages, too few for the channels-first loader)
    const uint32_t x_base = smem_u32(smem + pl.x), cbh_base = smem_u32(SPLIT ? cbhs : cbs), cbl_base = smem_u32(cbls);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nchunk = p.d >> 2;
    const float BIG = 1e30f;                    // score of a padded code (finite: keys stay ordered)
#ifdef TVQ_PROFILE_PHASES
    long long kt[8];
    kt[0] = clock64();
    if (tid == 0) atomicMin(&g_gt[0], gtimer());
#define TVQ_KT(i) kt[i] = clock64()
#else
#define TVQ_KT(i) do { } while (0)
#endif

    // ------------------------------------------------------------------ CTA prologue
    if ((smem_u32(smem) & 1023u) != 0) __trap();          // SWIZZLE_128B tiles need 1024-byte alignment
    if (tid == 0) {
        // "tile landed": one arrive.expect_tx by the TMA producer, or one arrive per loader warp (channels-first x)
        for (int s = 0; s < kUMaxStages; ++s) { mbar_init(bar_full + 8 * s, XCF ? 2 : 1); mbar_init(bar_empty + 8 * s, 4); }
        for (int s = 0; s < kUSlots; ++s) { mbar_init(bar_tfull + 8 * s, 1); mbar_init(bar_tempty + 8 * s, 4); }
        fence_mbar_init();
        if (!XCF) tma_prefetch_desc(&tmap_x);
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    // The first tiles of every CTA are fixed (tile = blockIdx.x * grab + j): their TMA loads are issued right here,
    // so that HBM latency overlaps the codebook set-up below; the rest are handed out by the global counter.
    const int num_tiles = p.num_tiles;                    // tiles of 64 rows
    int grab = (num_tiles / (int)gridDim.x) >> 1;
    grab = grab < 1 ? 1 : grab > stages ? stages : grab;  // gridDim.x <= num_tiles: every fixed tile exists
    if (!XCF && tid == 0) {
        for (int j = 0; j < grab; ++j) {
            const int tile = (int)blockIdx.x * grab + j;
            stage_tile[j] = tile;
            mbar_arrive_expect_tx(bar_full + 8 * j, (uint32_t)(kUM * DP * 4));
#pragma unroll
            for (int jj = 0; jj < NSLAB; ++jj)
                tma_load_2d(x_base + j * pl.stage_bytes + jj * SLAB_X, &tmap_x, bar_full + 8 * j, jj * 32, tile * kUM);
        }
    }
    {   // codebook -> UMMA B-operand layout (zero padded) and canonical |e|^2 (one warp per code, BIG for padding);
        // all loads of a thread are issued before the first use (one L2 round trip, not one per iteration)
        constexpr int NW = kUThreads / 32;
        constexpr int CB_IT = (KP * DPC + kUThreads - 1) / kUThreads, E_IT = (KP + NW - 1) / NW;
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* cb4 = reinterpret_cast<const float4*>(p.cb);
        float4 cv[CB_IT], ev[E_IT];
#pragma unroll
        for (int j = 0; j < CB_IT; ++j) {
            const int f = tid + j * kUThreads, row = f / DPC, c4 = f % DPC;
            cv[j] = (f < KP * DPC && row < p.k && c4 < nchunk) ? __ldcg(cb4 + (size_t)row * nchunk + c4) : z4;
        }
#pragma unroll
        for (int j = 0; j < E_IT; ++j) {
            const int c = warp + j * NW;
            ev[j] = (c < p.k && lane < nchunk) ? __ldcg(cb4 + (size_t)c * nchunk + lane) : z4;   // nchunk <= 32 here
        }
#pragma unroll
        for (int j = 0; j < CB_IT; ++j) {
            const int f = tid + j * kUThreads, row = f / DPC, c4 = f % DPC;
            if (f < KP * DPC) {
                // e = hi + lo exactly: hi keeps the 19 leading bits (sign, exponent, 10 mantissa bits: what kind::tf32 reads,
                // so its conversion is exact whether the tensor core truncates or rounds), lo is the exact remainder
                const float4 e = cv[j];
                float4 hi, lo;
                hi.x = __uint_as_float(__float_as_uint(e.x) & 0xFFFFE000u); lo.x = e.x - hi.x;
                hi.y = __uint_as_float(__float_as_uint(e.y) & 0xFFFFE000u); lo.y = e.y - hi.y;
                hi.z = __uint_as_float(__float_as_uint(e.z) & 0xFFFFE000u); lo.z = e.z - hi.z;
                hi.w = __uint_as_float(__float_as_uint(e.w) & 0xFFFFE000u); lo.w = e.w - hi.w;
                const int off = tile_off<KP>(row, c4);
                *reinterpret_cast<float4*>(cbs + off) = e;
                if (SPLIT) {
                    *reinterpret_cast<float4*>(cbhs + off) = hi;
                    *reinterpret_cast<float4*>(cbls + off) = lo;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < E_IT; ++j) {
            const int c = warp + j * NW;
            const double v = butterfly_sum(dot4(0.0, ev[j], ev[j]));
            if (c < KP && lane == 0) { e2s[c] = c < p.k ? __double2float_rn(v) : BIG; hist[c] = 0; }
        }
    }
    fence_proxy_async_smem();                             // generic-proxy writes -> visible to tcgen05.mma
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    TVQ_KT(1);
    float emax2 = 0.f;
    for (int c = 0; c < KP; ++c) emax2 = fmaxf(emax2, (c < p.k) ? e2s[c] : 0.f);
    const float emax = sqrtf(emax2) * 1.0001f;
    // |(s_a - s_b) - (d_a - d_b)| <= err_p * |x| * max|e| + err_s * (|x| + max|e|)^2.  The code words enter the tensor core
    // as hi + lo (two MMAs per K step; the pipe is ~8 % busy): hi is exact in tf32 and lo's conversion error is 2^-20 |e|,
    // so only x carries a tf32 error (within 2^-10 relative, truncated or rounded): 2^-10 * |x| |e| on a dot product, 2^-9 on
    // a score and 2^-8 * |x| * max|e| on a score difference (10 % slack: 4.4e-3) — HALF of what one MMA on the raw code words
    // gives, and half as many rows leave this level (round 2: 10 % -> 5 % of Gaussian rows, 40 % -> ~20 % of trained stage-1
    // latents).  err_s covers the fp32 accumulation (twice as many steps now), the roundings of both formulas and the 6 key
    // bits that carry the code (64 ulps).
    // (!SPLIT: both operands carry a tf32 error: 2^-7 * |x| * max|e|, 8.8e-3 with the slack.)
    const float err_p = (SPLIT ? 4.4e-3f : 8.8e-3f) * emax, err_s = 3e-5f;

    float loss = 0.f;
    unsigned counters = 0;                                // low 16 bits: re-scored rows, high: fp64 rows

    if (XCF && (warp == 0 || warp == kUWarpL2)) {
        // ============================================================ channels-first loaders (two warps, half a tile each)
        // Trip t issues iteration t (a tile, or an end marker once the tiles are exhausted) as one cp.async group and
        // completes iteration t - kXcfDepth.  Lane 0 of warp 0 is the only one that draws tiles; the id reaches the other
        // loader warp through stage_tile[] and a 64-thread named barrier.
        const int lw = warp == 0 ? 0 : 1;
        int issued = 0, ends = 0;
        for (int t = 0;; ++t) {
            if (ends < kUGroups) {
                const int s = t % stages;
                const uint32_t ph = (uint32_t)(t / stages) & 1u;
                mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                if (lw == 0 && lane == 0) {
                    int tile = -1;
                    if (ends == 0) {
                        tile = t < grab ? (int)blockIdx.x * grab + t : (int)gridDim.x * grab + (int)atomicAdd(&p.hdr->next_tile, 1u);
                        if (tile >= num_tiles) tile = -1;
                    }
                    stage_tile[s] = tile;
                }
                named_bar_sync(1, 64);
                const int tile = stage_tile[s];
                if (tile >= 0) xcf_issue_tile<DP>(p, x_base + s * pl.stage_bytes, tile, lw, lane);
                else ++ends;
                ++issued;
            }
            cp_async_commit();
            cp_async_wait_n<kXcfDepth>();
            const int a = t - kXcfDepth;
            if (a >= 0 && a < issued) {
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_full + 8 * (a % stages));
            }
            if (ends >= kUGroups && a >= issued - 1) break;
        }
    } else if (!XCF && warp == 0) {
        // ============================================================ TMA producer + dynamic tile scheduler
        // Tiles are handed out by a global counter (a CTA that starts late — e.g. behind the previous launch's
        // last CTA — simply takes fewer); the tile id travels with the stage.  After the last tile the producer
        // posts kUGroups end markers so that every epilogue group sees one.
        if (lane == 0) {
            int ends = 0;                                 // stages 0 .. grab-1 were filled in the prologue
            int s = grab % stages;
            uint32_t ph = (uint32_t)(grab / stages) & 1u;
            while (ends < kUGroups) {
                mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                int tile = -1;
                if (ends == 0) {
                    tile = (int)gridDim.x * grab + (int)atomicAdd(&p.hdr->next_tile, 1u);
                    if (tile >= num_tiles) tile = -1;
                }
                stage_tile[s] = tile;
                if (tile >= 0) {
                    mbar_arrive_expect_tx(bar_full + 8 * s, (uint32_t)(kUM * DP * 4));
#pragma unroll
                    for (int j = 0; j < NSLAB; ++j)
                        tma_load_2d(x_base + s * pl.stage_bytes + j * SLAB_X, &tmap_x, bar_full + 8 * s, j * 32, tile * kUM);
                } else {
                    mbar_arrive(bar_full + 8 * s);
                    ++ends;
                }
                if (++s == stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ============================================================ MMA issuer
        // The whole warp walks the loop and one elected lane issues: with warp-uniform control flow the descriptors live in
        // uniform registers and the 2 * d/8 MMAs of a tile go out back to back (under `if (lane == 0)` every MMA cost a
        // ~14-instruction elect / broadcast sequence: ~450 instructions of one thread between "tile landed" and "scores
        // ready").  Stage / slot indices and phases are carried, not divided out of the tile counter.
        {
            const bool leader = elect_one();
            constexpr uint32_t idesc = umma_idesc_tf32(kUM, KP);
            int s = 0, slot = 0;
            uint32_t ph = 0, sph = 0;
            for (;;) {
                mbar_wait(bar_full + 8 * s, ph);
                if (stage_tile[s] < 0) break;
                mbar_wait(bar_tempty + 8 * slot, sph ^ 1u);
                tc_fence_after();
                const uint32_t a0 = x_base + s * pl.stage_bytes;
                if (leader) {
#pragma unroll
                    for (int j = 0; j < NSLAB; ++j)
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const uint64_t adesc = umma_desc_sw128(a0 + j * SLAB_X + kk * 32);
                            umma_tf32(tmem_base + slot * KP, adesc, umma_desc_sw128(cbh_base + j * SLAB_CB + kk * 32), idesc, (j | kk) != 0);
                            if (SPLIT) umma_tf32(tmem_base + slot * KP, adesc, umma_desc_sw128(cbl_base + j * SLAB_CB + kk * 32), idesc, 1u);
                        }
                    umma_commit(bar_tfull + 8 * slot);
                }
                __syncwarp();
                if (++s == stages) { s = 0; ph ^= 1u; }
                if (++slot == kUSlots) { slot = 0; sph ^= 1u; }
            }
        }
    } else if (warp >= 2 && warp < kUWarpL2) {
        // ============================================================ epilogue warps
        const int g = (warp - 2) >> 2;                    // epilogue group: tiles g, g + kUGroups, ...
        const int quad = warp & 3;                        // TMEM lane quadrant this warp may access
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        const uint32_t acc_base = tmem_base + lane_base + kUSlots * KP + g * (4 * KP);
        const bool has_chunk = FULLD ? true : (lane < DPC && lane < nchunk);   // FULLD: d == DP == 128
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float4* keys = reinterpret_cast<float4*>(smem + pl.keys) + (warp - 2) * 16;   // this warp's 16 rows
        // per-lane address pieces of the swizzled tiles (tile_off with the lane's chunk folded in)
        const uint32_t l7s = (uint32_t)(lane & 7) << 4;                 // the lane's 16-byte chunk within a 128-byte row
        const int alane = has_chunk ? lane : 0;                         // lanes without a chunk read chunk 0 (discarded)
        const uint32_t cbw = smem_u32(cbs) + (uint32_t)(alane >> 3) * (KP * 128);
        if (TRAIN) {
            for (int c = 0; c < KP; ++c) tmem_st_x4(acc_base + 4 * c, z4);
            tmem_st_wait();
        }
#ifdef TVQ_PROFILE_PHASES
        long long ph_acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        long long ph_t = clock64();
#endif
        for (int it = g; ; it += kUGroups) {
            const int s = it % stages, slot = it % kUSlots;
            const uint32_t ph = (uint32_t)(it / stages) & 1u, sph = (uint32_t)(it / kUSlots) & 1u;
            const float* xt = reinterpret_cast<const float*>(smem + pl.x + s * pl.stage_bytes);

            mbar_wait(bar_full + 8 * s, ph);              // x tile landed (TMA writes visible) / end marker
            const int tile = stage_tile[s];
            if (tile < 0) break;
            const int64_t row0 = (int64_t)tile * kUM + quad * 16;     // first of this warp's 16 rows
            TVQ_PH(0);
#ifdef TVQ_PROFILE_PHASES
            const int tix = it / kUGroups;
            const bool trec = blockIdx.x == 0 && warp == 2 && lane == 0 && tix < 8;
            if (trec) g_tile_clk[tix][0] = (unsigned long long)(clock64() - kt[0]);
#endif
            mbar_wait(bar_tfull + 8 * slot, sph);         // scores ready
            tc_fence_after();
            TVQ_PH(1);
#ifdef TVQ_PROFILE_PHASES
            if (trec) g_tile_clk[tix][1] = (unsigned long long)(clock64() - kt[0]);
#endif
            // ---- 1. scan: lane l < 16 owns row l of the quadrant (M = 64 accumulator layout)
            {
                float t0 = BIG, t1 = BIG, t2 = BIG, t3 = BIG;
                float sc[KP];
#pragma unroll
                for (int c0 = 0; c0 < KP; c0 += 16) tmem_ld_x16(tmem_base + lane_base + (uint32_t)(slot * KP + c0), sc + c0);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty + 8 * slot);   // score slot may be overwritten
#pragma unroll
                for (int c = 0; c < KP; ++c) {
                    const float sv = fmaf(-2.f, sc[c], e2s[c]);
                    const float key = __uint_as_float((__float_as_uint(sv) & ~63u) | (unsigned)c);
                    top4_insert(key, t0, t1, t2, t3);
                }
                if (lane < 16) keys[lane] = make_float4(t0, t1, t2, t3);
                __syncwarp();
            }
            TVQ_PH(2);
#ifdef TVQ_PROFILE_PHASES
            if (trec) g_tile_clk[tix][2] = (unsigned long long)(clock64() - kt[0]);
#endif
            // ---- 2-4. apply: all lanes on one row, 4 rows in flight
            const int nvalid = (int)((p.n - row0) < 16 ? (p.n - row0) : 16);      // rows of this warp inside n
            if (nvalid == 16) {
                // shared-space byte addresses with the lane's chunk and swizzle folded in:
                //   x row r   : xw + r*128 + ((l7 ^ (r & 7)) << 4)      ((quad*16 + r) & 7 == r & 7)
                //   code row c: cbw + c*128 + ((l7 ^ (c & 7)) << 4)
                const uint32_t xw = smem_u32(xt) + (uint32_t)(alane >> 3) * (kUM * 128) + (uint32_t)quad * (16 * 128);
                float* qp = p.q ? p.q + (size_t)row0 * p.d + 4 * lane : nullptr;
                const size_t qstep = FULLD ? (size_t)DP : (size_t)p.d;
                int mycode = 0;                           // lane r (< 16) collects the code of row r
#pragma unroll 1
                for (int b = 0; b < 4; ++b) {
                    float4 xv[4], key[4];
                    float ss[4];
                    const uint32_t xb = xw + (uint32_t)b * 512u;
                    const uint32_t bsw = (uint32_t)(b & 1) << 6;           // (r & 7) << 4 = bsw | (u << 4)
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float4 ld = lds_v4(xb + u * 128 + ((l7s ^ (u << 4)) ^ bsw));
                        xv[u] = has_chunk ? ld : z4;
                        key[u] = keys[b * 4 + u];
                        ss[u] = fmaf(xv[u].x, xv[u].x, fmaf(xv[u].y, xv[u].y, fmaf(xv[u].z, xv[u].z, xv[u].w * xv[u].w)));
                    }
                    TVQ_PH(5);
                    {   // four warp-wide sums with 10 shuffles instead of 20: the first two butterfly steps also halve the
                        // number of values a lane carries (lanes with bit 4 set keep rows 2-3, bit 3 picks the row of
                        // the pair), three plain steps finish, four broadcasts hand every lane all four |x|^2 (the sums
                        // only feed the error bound, so their order is free)
                        const bool h4 = lane & 16, h3 = lane & 8;
                        const float k0 = h4 ? ss[2] : ss[0], k1 = h4 ? ss[3] : ss[1];
                        const float g0 = h4 ? ss[0] : ss[2], g1 = h4 ? ss[1] : ss[3];
                        const float r0 = k0 + __shfl_xor_sync(0xffffffffu, g0, 16);
                        const float r1 = k1 + __shfl_xor_sync(0xffffffffu, g1, 16);
                        float t = (h3 ? r1 : r0) + __shfl_xor_sync(0xffffffffu, h3 ? r0 : r1, 8);
                        t += __shfl_xor_sync(0xffffffffu, t, 4);
                        t += __shfl_xor_sync(0xffffffffu, t, 2);
                        t += __shfl_xor_sync(0xffffffffu, t, 1);
                        ss[0] = __shfl_sync(0xffffffffu, t, 0);     // lane (bit4, bit3) = (0,0) -> row 0, (0,1) -> 1, (1,0) -> 2, (1,1) -> 3
                        ss[1] = __shfl_sync(0xffffffffu, t, 8);
                        ss[2] = __shfl_sync(0xffffffffu, t, 16);
                        ss[3] = __shfl_sync(0xffffffffu, t, 24);
                    }
                    TVQ_PH(6);
                    int cd[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {          // (a) decisions
                        const float xn = sqrt_approx(ss[u]) * 1.0001f;
                        const float bnd = xn + emax;
                        const float bnd2 = bnd * bnd;
                        const float lim = fmaf(err_p, xn, fmaf(err_s, bnd2, key[u].x));
                        int code = (int)(__float_as_uint(key[u].x) & 63u);
                        if (!(key[u].y > lim)) {
                            // second best inside the bound (or non-finite scores: the comparison is
                            // false for NaN): cascade re-score
                            const int nc = 1 + (key[u].y <= lim) + (key[u].z <= lim) + (key[u].w <= lim);
                            const bool finite = key[u].x < BIG && bnd2 < BIG;
                            const int r = resolve_row<KP>(xv[u], bnd2, __float_as_uint(key[u].x) & 63u,
                                                          __float_as_uint(key[u].y) & 63u, __float_as_uint(key[u].z) & 63u, nc,
                                                          nc == 4 || !finite, cbs, e2s, p.k, has_chunk, lane);
                            counters += 1u + ((unsigned)(r >> 16) << 16);
                            code = r & 0xffff;
                            code = code < p.k ? code : 0;  // non-finite rows: any code, but a valid one
                        }
                        cd[u] = code;
                    }
                    TVQ_PH(9);
                    if (p.q != nullptr || TRAIN) {
                        float4 ev[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {      // (b) gather from the shared-memory codebook
                            const uint32_t c = (uint32_t)cd[u];
                            const float4 ld = lds_v4(cbw + (c << 7) + (((c << 4) ^ l7s) & 0x70u));
                            ev[u] = has_chunk ? ld : z4;
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {      // (c) straight-through, loss, store
                            float4 o = ev[u];
                            if (TRAIN) {
                                // x + (e - x): two rounded fp32 ops, never contracted; the loss is taken
                                // on that rounded tensor, as F.mse_loss(quantize.detach(), x) does
                                o.x = __fadd_rn(xv[u].x, __fsub_rn(ev[u].x, xv[u].x));
                                o.y = __fadd_rn(xv[u].y, __fsub_rn(ev[u].y, xv[u].y));
                                o.z = __fadd_rn(xv[u].z, __fsub_rn(ev[u].z, xv[u].z));
                                o.w = __fadd_rn(xv[u].w, __fsub_rn(ev[u].w, xv[u].w));
                                const float dx = __fsub_rn(o.x, xv[u].x), dy = __fsub_rn(o.y, xv[u].y);
                                const float dz = __fsub_rn(o.z, xv[u].z), dw = __fsub_rn(o.w, xv[u].w);
                                loss = fmaf(dx, dx, loss);
                                loss = fmaf(dy, dy, loss);
                                loss = fmaf(dz, dz, loss);
                                loss = fmaf(dw, dw, loss);
                            }
                            if (qp != nullptr && has_chunk) {
                                if (p.q_hw > 0) {
                                    // channels-first output: park q in the x tile (this lane's own chunk, already in
                                    // registers) and store the warp's 16 rows transposed once they are all done
                                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(xb + u * 128 + ((l7s ^ (u << 4)) ^ bsw)),
                                                 "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
                                } else {
                                    st_stream_v4(qp + u * qstep, o);
                                }
                            }
                        }
                        if (qp != nullptr) qp += 4 * qstep;
                    }
                    {   // lane 4b+u remembers the code of row 4b+u (idx store and counts happen once per tile)
                        const unsigned pk = (unsigned)cd[0] | ((unsigned)cd[1] << 8) | ((unsigned)cd[2] << 16) | ((unsigned)cd[3] << 24);
                        const int mine = (int)((pk >> ((lane & 3) << 3)) & 0xffu);
                        mycode = ((lane >> 2) == b) ? mine : mycode;
                    }
                    TVQ_PH(7);
                    if (TRAIN) {
                        // ONE tensor-memory read-modify-write per row; rows of the batch that share a
                        // code (a warp-uniform condition) take the sequential form.  (Folding such rows into one
                        // read-modify-write was measured and is slower: the predicated loads / stores cost the
                        // common path more than the sequential form costs the rare one.)
                        const bool dup = cd[1] == cd[0] || cd[2] == cd[0] || cd[2] == cd[1] || cd[3] == cd[0] ||
                                         cd[3] == cd[1] || cd[3] == cd[2];
                        __syncwarp();
                        tmem_st_wait();                   // the previous batch's stores have landed
                        if (!dup) {
                            float4 a[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) a[u] = tmem_ld_x4(acc_base + 4 * cd[u]);
                            tmem_ld_wait();
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                a[u].x += xv[u].x; a[u].y += xv[u].y; a[u].z += xv[u].z; a[u].w += xv[u].w;
                                tmem_st_x4(acc_base + 4 * cd[u], a[u]);
                            }
                        } else {
#pragma unroll 1
                            for (int u = 0; u < 4; ++u) {
                                const int c = u == 0 ? cd[0] : u == 1 ? cd[1] : u == 2 ? cd[2] : cd[3];
                                const float4 xa = u == 0 ? xv[0] : u == 1 ? xv[1] : u == 2 ? xv[2] : xv[3];
                                float4 a = tmem_ld_x4(acc_base + 4 * c);
                                tmem_ld_wait();
                                a.x += xa.x; a.y += xa.y; a.z += xa.z; a.w += xa.w;
                                tmem_st_x4(acc_base + 4 * c, a);
                                tmem_st_wait();
                            }
                        }
                    }
                    TVQ_PH(8);
                }
                if (lane < 16) {                            // idx + counts for the warp's 16 rows
                    atomicAdd(hist + mycode, 1);
                    p.idx[row0 + lane] = (int64_t)mycode;
                }
                if (p.q != nullptr && p.q_hw > 0) {
                    const bool tile_full = (int64_t)tile * kUM + kUM <= p.n;    // (the same for the four warps of the group)
                    if (tile_full) {
                        // transposed store by the whole group: once its four warps have parked their rows, each stores a
                        // quarter of the channels for ALL 64 rows of the tile — lane = latent, so one instruction writes 32
                        // consecutive positions of 'b c (h w)' (128-byte runs; 64-byte runs per warp cost 20-40 % more time
                        // at large n: partial sectors)
                        named_bar_sync(2 + g, 128);
                        const int cq = (p.d + 3) >> 2;                          // channels per warp
#pragma unroll
                        for (int hrow = 0; hrow < 2; ++hrow) {
                            const int r = 32 * hrow + lane;
                            const int64_t grow = (int64_t)tile * kUM + r;
                            const int64_t bi = grow / p.q_hw, hw = grow - bi * p.q_hw;
                            float* qc = p.q + bi * (int64_t)p.d * p.q_hw + hw;
                            const float* trow = xt + r * 32;                      // row inside a 32-float slab
                            const int rsw = r & 7;
                            const int c1 = (quad + 1) * cq < p.d ? (quad + 1) * cq : p.d;
#pragma unroll 4
                            for (int c = quad * cq; c < c1; ++c) {
                                const float v = trow[(c >> 5) * (kUM * 32) + ((((c >> 2) & 7) ^ rsw) << 2) + (c & 3)];
                                qc[(int64_t)c * p.q_hw] = v;
                            }
                        }
                        __syncwarp();
                    } else {
                        // (a full quadrant of the last, partial tile) the warp's own 16 rows: lane = (row, channel parity)
                        __syncwarp();
                        const int r = lane & 15, par = lane >> 4;
                        const int64_t grow = row0 + r;
                        const int64_t bi = grow / p.q_hw, hw = grow - bi * p.q_hw;
                        float* qc = p.q + bi * (int64_t)p.d * p.q_hw + hw;
                        const float* trow = xt + (quad * 16 + r) * 32;             // row inside a 32-float slab
                        const int rsw = (quad * 16 + r) & 7;
#pragma unroll 4
                        for (int cc = 0; cc < (p.d >> 1); ++cc) {
                            const int c = 2 * cc + par;
                            const float v = trow[(c >> 5) * (kUM * 32) + ((((c >> 2) & 7) ^ rsw) << 2) + (c & 3)];
                            qc[(int64_t)c * p.q_hw] = v;
                        }
                        __syncwarp();
                    }
                }
            } else {
                for (int r = 0; r < nvalid; ++r) {        // last, partial tile: one fully checked row at a time
                    const RowResult rr = row_generic<DP, KP, TRAIN>(p.q, p.idx, p.d, p.k, xt, cbs, e2s, hist, keys[r], quad * 16 + r,
                                                                    row0 + r, has_chunk, lane, emax, err_p, err_s, acc_base, p.q_hw);
                    loss += rr.loss;
                    counters += rr.counters;
                }
            }
            TVQ_PH(3);
#ifdef TVQ_PROFILE_PHASES
            if (trec) g_tile_clk[tix][3] = (unsigned long long)(clock64() - kt[0]);
#endif
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + 8 * s);   // this warp is done with the stage
            TVQ_PH(4);
        }
        if (TRAIN) tmem_st_wait();
#ifdef TVQ_PROFILE_PHASES
        if (blockIdx.x == 0 && lane == 0 && quad == 0 && g < 2)
        {
            for (int i = 0; i < 8; ++i) g_phase_clk[g][i] = (unsigned long long)ph_acc[i];
            for (int i = 0; i < 10; ++i) g_phase_all[g][i] = (unsigned long long)ph_acc[i];
        }
#endif
    }

    // ------------------------------------------------------------------ teardown and flush
    TVQ_KT(2);
    tc_fence_before();
    __syncthreads();                                      // all tiles consumed: the stages are free
    tc_fence_after();
    TVQ_KT(3);
    // Every epilogue warp adds its TMEM accumulators into the (now idle) tile ring: one [4 quadrants][KP][32 lanes]
    // float4 area per group when the ring is large enough for all groups at once (d = 128: one pass), otherwise
    // `par` groups per pass; then the CTA's total goes out with one red.global.add.v4 per 16 bytes.
    constexpr int DUMP_CELLS = 4 * KP * 32;
    float4* dump = reinterpret_cast<float4*>(smem + pl.x);
    int par = (stages * pl.stage_bytes) / (DUMP_CELLS * 16);
    par = par > kUGroups ? kUGroups : par;                // >= 1 (checked by the host)
    if (TRAIN) {
        for (int g0 = 0; g0 < kUGroups; g0 += par) {
            const int g = (warp >= 2 && warp < kUWarpL2) ? (warp - 2) >> 2 : -1;
            if (g >= g0 && g < g0 + par) {
                const int quad = warp & 3;
                const uint32_t acc_base = tmem_base + ((uint32_t)(quad * 32) << 16) + kUSlots * KP + g * (4 * KP);
                float4* area = dump + (g - g0) * DUMP_CELLS + quad * KP * 32 + lane;
                for (int c0 = 0; c0 < KP; c0 += 8) {
                    float4 v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = tmem_ld_x4(acc_base + 4 * (c0 + j));
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4* cell = area + (c0 + j) * 32;
                        if (g0 > 0) { const float4 o = *cell; v[j].x += o.x; v[j].y += o.y; v[j].z += o.z; v[j].w += o.w; }
                        *cell = v[j];
                    }
                }
            }
            tc_fence_before();
            __syncthreads();
        }
    }
    TVQ_KT(4);
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
    if (TRAIN) {
        float* esum = p.stats + ((p.k + 3) & ~3);
        for (int f = tid; f < KP * 32; f += kUThreads) {
            const int c = f >> 5, l = f & 31;
            if (c < p.k && l < nchunk && l < DPC) {
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int w = 0; w < 4 * par; ++w) {       // (area, quadrant) pairs: area stride = 4 * KP * 32
                    const float4 b = dump[(w * KP + c) * 32 + l];
                    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
                }
                if (a.x != 0.f || a.y != 0.f || a.z != 0.f || a.w != 0.f) red_add_v4(esum + (size_t)c * p.d + 4 * l, a);
            }
        }
    }
    for (int c = tid; c < p.k; c += kUThreads) {
        const int v = hist[c];
        if (v) atomicAdd(p.stats + c, (float)v);
    }
    if (lane == 0 && counters) {
        atomicAdd(&p.hdr->n_rescored, counters & 0xffffu);
        if (counters >> 16) atomicAdd(&p.hdr->n_exact, counters >> 16);
    }
    TVQ_KT(5);
    if (TRAIN) {
        double t = block_sum((double)loss, red);
        if (tid == 0) atomicAdd(&p.hdr->loss_sum, t);
    }
    TVQ_KT(6);
    finish_resident<KP, TRAIN>(p, cbs, misc);
    TVQ_KT(7);
#ifdef TVQ_PROFILE_PHASES
    if (blockIdx.x == 0 && (tid == 0 || tid == 64))
        for (int i = 0; i < 8; ++i) g_phase_clk[tid == 0 ? 0 : 1][8 + i] = (unsigned long long)(kt[i] - kt[0]);
    if (tid == 0) {
        atomicMax(&g_gt[1], gtimer());
        atomicMax(&g_gt[2], (unsigned long long)(kt[7] - kt[0]));
        atomicMax(&g_gt[3], (unsigned long long)(kt[3] - kt[0]));
    }
#endif
}

}  // namespace tvq
