// tvq_fwd_umma.cuh — fused VQ forward with tcgen05 scoring, codebook resident in shared memory
// (k <= 64 codes, d <= 128).  This is the path BASELINE configs[0..1,3,4] run (k = 32, d = 128).
//
// One persistent CTA per SM, 10 warps, warp-specialised:
//   warp 0      TMA producer: x tiles of 64 latents x d fp32 -> ring of shared-memory stages
//               (cp.async.bulk.tensor, SWIZZLE_128B, zero-fill past n / past d)
//   warp 1      MMA issuer: per tile d/8 tcgen05.mma.kind::tf32 (M=64, N=KP) straight from the fp32
//               tiles (no conversion pass) into one of 4 TMEM accumulator slots; tcgen05.commit
//   warps 2-5   epilogue group 0, warps 6-9 epilogue group 1: the groups take alternate tiles and
//               never synchronise with each other, so one group's latencies hide behind the other's.
// Per tile an epilogue group
//   1. reads the 64 x KP approximate dot products from TMEM (tcgen05.ld), forms scores
//      e2 - 2 x.e and keeps, per row, every code within the RIGOROUS tf32 error bound of the row
//      minimum;
//   2. decides rows with more than one candidate by the canonical fp64 re-score
//      (tvq_common.cuh, mirrored by oracle/vq_canon.c) — so the indices are exactly those of the
//      SIMT path and of the C oracle, whatever the tensor-core rounding did;
//   3. writes idx, gathers the code word, forms the straight-through output and the commitment
//      loss partial, streams q_st out (st.global.cs);
//   4. buckets the tile's rows by code and adds them, column-owner style, into REGISTER
//      accumulators (no floating-point atomics until the CTA's single flush);
//   5. releases the stage to the producer.
// Algorithmic HBM traffic: read x once, write q once, write idx: 8d + 8 bytes per latent.
#pragma once
#include <cuda.h>

#include "tvq_common.cuh"
#include "tvq_fwd_simt.cuh"
#include "tvq_sm100.cuh"

namespace tvq {

constexpr int kUM = 64;                 // latents per UMMA tile (M)
constexpr int kUGroupThreads = 128;     // one epilogue group = 4 warps = the 4 TMEM lane quadrants
constexpr int kUThreads = 64 + 2 * kUGroupThreads;
constexpr int kUSlots = 4;              // TMEM accumulator slots
constexpr int kUMaxStages = 8;

struct UmmaPlan {
    int stages, stage_bytes;
    int x, cb, e2s, grp, grp_stride, red, misc, bars, tmem, total;
    // per-group block: sidx[64] order[64] xn2[64] start[KP+4] cntw[2*KP] hist[KP]
};
__host__ __device__ inline UmmaPlan make_umma_plan(int dp, int kp, int stages) {
    UmmaPlan u;
    u.stages = stages;
    u.stage_bytes = kUM * dp * 4;
    int o = 0;
    u.x = o;    o += stages * u.stage_bytes;
    u.cb = o;   o += kp * dp * 4;
    u.e2s = o;  o += kp * 4;
    u.grp = o;  u.grp_stride = (64 + 64 + 64 + (kp + 4) + 2 * kp + kp) * 4;
    o += 2 * u.grp_stride;
    o = (o + 15) & ~15;
    u.red = o;  o += 16 * 8;
    u.misc = o; o += 16 * 4;
    u.bars = o; o += (2 * kUMaxStages + 2 * kUSlots) * 8;
    u.tmem = o; o += 16;
    u.total = o;
    return u;
}

template <int DP, int KP, bool TRAIN>
__global__ void __launch_bounds__(kUThreads, 1) fwd_umma_kernel(const __grid_constant__ CUtensorMap tmap_x, const FwdParams p,
                                                               const int stages) {
    using namespace sm100;
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int DPC = DP / 4;                 // 16-byte chunks per padded row
    constexpr int NSLAB = DP / 32;              // 128-byte K slabs per tile
    constexpr int SLAB_X = kUM * 128;           // bytes of one x slab
    constexpr int SLAB_CB = KP * 128;           // bytes of one codebook slab
    constexpr int NSUB = kUGroupThreads / DP > 0 ? kUGroupThreads / DP : 1;   // threads sharing one column
    constexpr int NACC = KP / NSUB;             // register accumulators per thread
    constexpr uint32_t TMEM_COLS = (kUSlots * KP) <= 32 ? 32 : (kUSlots * KP) <= 64 ? 64 : (kUSlots * KP) <= 128 ? 128 : 256;
    static_assert(DP == 64 || DP == 128, "resident-codebook path: d padded to 64 or 128");
    static_assert(KP == 16 || KP == 32 || KP == 64, "resident-codebook path: k padded to 16, 32 or 64");

    const UmmaPlan pl = make_umma_plan(DP, KP, stages);
    float* cbs = reinterpret_cast<float*>(smem + pl.cb);
    float* e2s = reinterpret_cast<float*>(smem + pl.e2s);
    double* red = reinterpret_cast<double*>(smem + pl.red);
    int* misc = reinterpret_cast<int*>(smem + pl.misc);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + pl.bars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + pl.tmem);
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * kUMaxStages;
    const uint32_t bar_tfull = bar_empty + 8 * kUMaxStages, bar_tempty = bar_tfull + 8 * kUSlots;
    const uint32_t x_base = smem_u32(smem + pl.x), cb_base = smem_u32(cbs);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nchunk = p.d >> 2;
    const float INF = __int_as_float(0x7f800000);

    // ------------------------------------------------------------------ CTA prologue
    if ((smem_u32(smem) & 1023u) != 0) __trap();          // SWIZZLE_128B tiles need 1024-byte alignment
    for (int f = tid; f < KP * DPC; f += kUThreads) {     // codebook -> UMMA B-operand layout (zero padded)
        const int row = f / DPC, c4 = f % DPC;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < p.k && c4 < nchunk) v = __ldg(reinterpret_cast<const float4*>(p.cb + (size_t)row * p.d) + c4);
        *reinterpret_cast<float4*>(cbs + tile_off<KP>(row, c4)) = v;
    }
    for (int c = tid; c < KP; c += kUThreads) e2s[c] = (c < p.k) ? __ldg(p.e2 + c) : INF;
    for (int g = 0; g < 2; ++g) {                         // per-group histograms
        int* hist = reinterpret_cast<int*>(smem + pl.grp + g * pl.grp_stride) + 64 + 64 + 64 + (KP + 4) + 2 * KP;
        for (int c = tid; c < KP; c += kUThreads) hist[c] = 0;
    }
    fence_proxy_async_smem();                             // generic-proxy writes -> visible to tcgen05.mma
    if (tid == 0) {
        for (int s = 0; s < kUMaxStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int s = 0; s < kUSlots; ++s) { mbar_init(bar_tfull + 8 * s, 1); mbar_init(bar_tempty + 8 * s, 4); }
        fence_mbar_init();
        tma_prefetch_desc(&tmap_x);
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    float emax2 = 0.f;
    for (int c = 0; c < KP; ++c) emax2 = fmaxf(emax2, (c < p.k) ? e2s[c] : 0.f);
    const float emax = sqrtf(emax2) * 1.0001f;
    // |(s_a - s_b) - (d_a - d_b)| <= err_c * (|x| + max|e|)^2 for tf32 operands (each within 2^-10
    // relative, truncated or rounded) and fp32 accumulation: 2^-9 with 10 % slack + fp32 terms.
    const float err_c = 2.2e-3f;

    const int num_tiles = p.num_tiles;                    // tiles of 64 rows
    ApplyState st;
    st.loss = 0.f;
    float acc[NACC];
#pragma unroll
    for (int j = 0; j < NACC; ++j) acc[j] = 0.f;

    if (warp == 0) {
        // ============================================================ TMA producer
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int s = it % stages;
                const uint32_t ph = (uint32_t)(it / stages) & 1u;
                mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                mbar_arrive_expect_tx(bar_full + 8 * s, (uint32_t)(kUM * DP * 4));
#pragma unroll
                for (int j = 0; j < NSLAB; ++j)
                    tma_load_2d(x_base + s * pl.stage_bytes + j * SLAB_X, &tmap_x, bar_full + 8 * s, j * 32, tile * kUM);
            }
        }
    } else if (warp == 1) {
        // ============================================================ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(kUM, KP);
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int s = it % stages, slot = it % kUSlots;
                const uint32_t ph = (uint32_t)(it / stages) & 1u, sph = (uint32_t)(it / kUSlots) & 1u;
                mbar_wait(bar_full + 8 * s, ph);
                mbar_wait(bar_tempty + 8 * slot, sph ^ 1u);
                tc_fence_after();
                const uint32_t a0 = x_base + s * pl.stage_bytes;
#pragma unroll
                for (int j = 0; j < NSLAB; ++j)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_tf32(tmem_base + slot * KP, umma_desc_sw128(a0 + j * SLAB_X + kk * 32),
                                  umma_desc_sw128(cb_base + j * SLAB_CB + kk * 32), idesc, (j | kk) != 0);
                umma_commit(bar_tfull + 8 * slot);
            }
        }
    } else {
        // ============================================================ epilogue groups
        const int g = (warp - 2) >> 2;                    // group 0 / 1
        const int lt = tid - 64 - g * kUGroupThreads;     // 0..127 within the group
        const int wg = lt >> 5;                           // warp within the group
        const int quad = warp & 3;                        // TMEM lane quadrant this warp may read
        int* gi = reinterpret_cast<int*>(smem + pl.grp + g * pl.grp_stride);
        int* sidx = gi;
        int* order = gi + 64;
        float* xn2 = reinterpret_cast<float*>(gi + 128);
        int* start = gi + 192;
        int* cntw = start + (KP + 4);
        int* hist = cntw + 2 * KP;
        const uint32_t bar_id = 1 + g;

        for (int it = g; ; it += 2) {
            const int tile = blockIdx.x + it * gridDim.x;
            if (tile >= num_tiles) break;
            const int s = it % stages, slot = it % kUSlots;
            const uint32_t ph = (uint32_t)(it / stages) & 1u, sph = (uint32_t)(it / kUSlots) & 1u;
            const int64_t row0 = (int64_t)tile * kUM;
            const float* xt = reinterpret_cast<const float*>(smem + pl.x + s * pl.stage_bytes);

            mbar_wait(bar_full + 8 * s, ph);              // x tile landed (TMA writes visible)
            // ---- row norms of this warp's 16 rows (overlaps the MMA)
            for (int r = 0; r < 16; ++r) {
                const int row = quad * 16 + r;
                float ss = 0.f;
                for (int c = lane; c < DPC; c += 32) {
                    float4 v = *reinterpret_cast<const float4*>(xt + tile_off<kUM>(row, c));
                    ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
                }
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
                if (lane == 0) xn2[row] = ss;
            }
            __syncwarp();
            // ---- scores from TMEM: lane l < 16 owns row quad*16 + l (M = 64 accumulator layout)
            mbar_wait(bar_tfull + 8 * slot, sph);
            tc_fence_after();
            float sc[KP];
#pragma unroll
            for (int c0 = 0; c0 < KP; c0 += 16)
                tmem_ld_x16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(slot * KP + c0), sc + c0);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * slot);   // accumulator slot may be overwritten
            const int row = quad * 16 + (lane & 15);
            const bool active = lane < 16 && row0 + row < p.n;
            float m = INF;
#pragma unroll
            for (int c = 0; c < KP; ++c) {
                sc[c] = fmaf(-2.f, sc[c], e2s[c]);
                m = fminf(m, sc[c]);
            }
            const float b = sqrtf(xn2[row]) * 1.0001f + emax;
            const float lim = m + err_c * b * b;
            int cand[4] = {0, 0, 0, 0};
            int ncand = 0;
#pragma unroll
            for (int c = 0; c < KP; ++c) {
                if (sc[c] <= lim) {
                    if (ncand == 0) cand[0] = c;
                    else if (ncand == 1) cand[1] = c;
                    else if (ncand == 2) cand[2] = c;
                    else if (ncand == 3) cand[3] = c;
                    ++ncand;
                }
            }
            if (!(m < INF)) ncand = 5;                    // NaN / Inf rows: let the exact scan decide
            int code = cand[0];
            // ---- canonical fp64 re-score of rows with more than one candidate (warp-cooperative)
            unsigned need = __ballot_sync(0xffffffffu, active && ncand > 1);
            unsigned nres = __popc(need), nfull = 0;
            while (need) {
                const int src = __ffs(need) - 1;
                need &= need - 1;
                const int r_row = __shfl_sync(0xffffffffu, row, src);
                const int r_n = __shfl_sync(0xffffffffu, ncand, src);
                int r_c[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) r_c[j] = __shfl_sync(0xffffffffu, cand[j], src);
                int best;
                if (r_n > 4) { best = canon_scan_row<DP, kUM>(xt, r_row, p.cb, p.e2, p.k, p.d, lane); ++nfull; }
                else best = canon_pick<DP, kUM>(xt, r_row, p.cb, p.e2, p.d, lane, r_c, r_n);
                if (lane == src) code = best;
            }
            if (lane == 0 && nres) { atomicAdd(&p.hdr->n_rescored, nres); if (nfull) atomicAdd(&p.hdr->n_exact, nfull); }
            if (lane < 16) sidx[row] = active ? code : 0;
            if (active) p.idx[row0 + row] = (int64_t)code;
            named_bar_sync(bar_id, kUGroupThreads);
            // ---- gather, straight-through, loss, q_st out
            apply_rows<DP, TRAIN, kUM>(p, xt, sidx, hist, row0, st, wg, 4);
            // ---- EMA statistics into register accumulators
            if (TRAIN) {
                for (int i = lt; i < 2 * KP; i += kUGroupThreads) cntw[i] = 0;
                named_bar_sync(bar_id, kUGroupThreads);
                int mycode = -1, rank = 0;
                if (lt < kUM) {
                    mycode = (row0 + lt < p.n) ? sidx[lt] : -1;
                    const unsigned mm = __match_any_sync(0xffffffffu, mycode);
                    rank = __popc(mm & lanemask_lt());
                    if (mycode >= 0 && rank == 0) cntw[wg * KP + mycode] = __popc(mm);
                }
                named_bar_sync(bar_id, kUGroupThreads);
                if (wg == 0) {
                    constexpr int PER = KP / 32 > 0 ? KP / 32 : 1;
                    int v[PER], tot = 0;
#pragma unroll
                    for (int j = 0; j < PER; ++j) {
                        const int c = lane * PER + j;
                        v[j] = (c < KP) ? cntw[c] + cntw[KP + c] : 0;
                        tot += v[j];
                    }
                    int incl = tot;
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        int t = __shfl_up_sync(0xffffffffu, incl, off);
                        if (lane >= off) incl += t;
                    }
                    int run = incl - tot;
#pragma unroll
                    for (int j = 0; j < PER; ++j) {
                        const int c = lane * PER + j;
                        if (c < KP) start[c] = run;
                        run += v[j];
                    }
                    if (lane == 31) start[KP] = incl;
                }
                named_bar_sync(bar_id, kUGroupThreads);
                if (lt < kUM && mycode >= 0) order[start[mycode] + rank + (wg == 1 ? cntw[mycode] : 0)] = lt;
                named_bar_sync(bar_id, kUGroupThreads);
                const int col = lt % DP, sub = lt / DP;
                const int xc4 = col >> 2, xo = col & 3;
#pragma unroll
                for (int j = 0; j < NACC; ++j) {
                    const int c = j * NSUB + sub;
                    const int beg = start[c], end = start[c + 1];
                    float a = 0.f;
                    for (int i = beg; i < end; ++i) a += xt[tile_off<kUM>(order[i], xc4) + xo];
                    acc[j] += a;
                    if (col == 0 && end > beg) hist[c] += end - beg;
                }
            }
            named_bar_sync(bar_id, kUGroupThreads);       // every read of the stage is done
            if (lt == 0) mbar_arrive(bar_empty + 8 * s);
        }
    }

    // ------------------------------------------------------------------ teardown and flush
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
    float* scratch = reinterpret_cast<float*>(smem + pl.x);          // stage 0 is free now: [KP][DP] sums
    if (TRAIN) {
        if (warp >= 2 && warp < 6) {
            const int lt = tid - 64, col = lt % DP, sub = lt / DP;
#pragma unroll
            for (int j = 0; j < NACC; ++j) scratch[(j * NSUB + sub) * DP + col] = acc[j];
        }
        __syncthreads();
        if (warp >= 6) {
            const int lt = tid - 64 - kUGroupThreads, col = lt % DP, sub = lt / DP;
#pragma unroll
            for (int j = 0; j < NACC; ++j) scratch[(j * NSUB + sub) * DP + col] += acc[j];
        }
        __syncthreads();
        float* esum = p.stats + ((p.k + 3) & ~3);
        for (int f = tid; f < KP * DPC; f += kUThreads) {
            const int c = f / DPC, c4 = f % DPC;
            if (c < p.k && c4 < nchunk) {
                const float4 v = *reinterpret_cast<const float4*>(scratch + c * DP + 4 * c4);
                if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) red_add_v4(esum + (size_t)c * p.d + 4 * c4, v);
            }
        }
    }
    {
        const int* h0 = reinterpret_cast<const int*>(smem + pl.grp) + 64 + 64 + 64 + (KP + 4) + 2 * KP;
        const int* h1 = reinterpret_cast<const int*>(smem + pl.grp + pl.grp_stride) + 64 + 64 + 64 + (KP + 4) + 2 * KP;
        for (int c = tid; c < p.k; c += kUThreads) {
            const int v = h0[c] + h1[c];
            if (v) atomicAdd(p.stats + c, (float)v);
        }
    }
    if (TRAIN) {
        double t = block_sum((double)st.loss, red);
        if (tid == 0) atomicAdd(&p.hdr->loss_sum, t);
    }
    finish_ticket<TRAIN>(p, red, misc);
}

}  // namespace tvq
