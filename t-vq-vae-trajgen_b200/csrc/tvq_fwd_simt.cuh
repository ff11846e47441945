// tvq_fwd_simt.cuh — fused VQ forward, CUDA-core scoring (any k, d <= 256, d % 4 == 0).
//
// One persistent CTA walks tiles of 128 latents.  Per tile:
//   L  load the x tile into shared memory (cp.async, UMMA-compatible swizzled layout)
//   S  score it against the codebook in fp32 FMAs (register-tiled 4x4), keeping per row the best
//      and second-best approximate score
//   R  rows whose top-2 gap is inside the rigorous fp32 error bound are decided by the
//      canonical fp64 scan (tvq_common.cuh) — the result is therefore the canonical argmin for
//      EVERY row, whatever the scoring precision
//   A  write idx; gather the code word, straight-through output, commitment-loss partial sum
//      (vq.py:225, :358-364); accumulate the EMA statistics counts / embed_sum (vq.py:228-233)
// and after the last tile the CTA flushes its statistics; the last CTA to finish turns the
// totals into the two scalars (commit loss, perplexity vq.py:246-247).
//
// This path is the generic fallback and the bring-up reference for the tcgen05 path
// (tvq_fwd_umma.cuh), which replaces phases L and S only.
#pragma once
#include "tvq_common.cuh"

namespace tvq {

constexpr int kBN = 32;  // codes per SIMT code tile

enum StatsMode : int { kStatsNone = 0, kStatsSmall = 1, kStatsLarge = 2 };

struct FwdParams {
    const float* x;
    const float* cb;
    int64_t n;
    int k, d;
    int64_t* idx;
    float* q;         // may be null
    float* stats;     // [roundup(k,4) + k*d]: counts, then embed_sum (16-byte aligned)
    float* scalars;   // [TVQ_NUM_SCALARS]
    float commitment_weight;
    WsHeader* hdr;
    const float* e2;  // [k] canonical |e|^2 (prep kernel)
    int num_tiles;
    int stats_mode;   // StatsMode
    int use_hist;     // shared-memory count histogram (k <= 2048)
    int exact;        // decide every row with the canonical scan
    int given_idx;    // skip scoring: idx[] already holds the codes (stochastic / dropout branches)
    float* commit_out;    // optional 1-element copies of scalars[0] / scalars[2] (separate autograd outputs)
    float* weighted_out;
    // fused EMA update by the last CTA (tvq_train_step): buffers updated in place, statistics consumed
    int fuse_ema;
    float* cluster_size;
    float* embed_avg;
    float* embed;         // == cb (written only after every CTA has finished reading it)
    float* embed_prev;    // optional: receives the pre-update codebook
    float decay, one_minus_decay, eps, k_eps;
    // data-parallel fused step: exchange buffers of every rank (tvq_aux.cuh::ema_dp_kernel layout); world <= 1: local
    void* const* peers;
    int dp_rank, dp_world;
    int dp_defer;         // data-parallel: the last CTA only PUBLISHES this rank's statistics (no wait, no EMA update); the update
                          // runs later in tvq_ema_finalize_dp, off the critical path (tvq_hint_defer_exchange)
    unsigned long long dp_timeout_ns;   // give up waiting for a peer after this long (0 = never); tvq_common.cuh::wait_flag_sys
    // q layout: 0 = [n, d] row-major; hw > 0 = channels-first [n / hw, d, hw] (the 'b c (h w)' layout of the caller,
    // utils/train_utils.py:349): resident-codebook tcgen05 kernel only
    int q_hw;
    // x layout: 0 = [n, d] row-major; hw > 0 = channels-first [n / hw, d, hw] (== q_hw when q is written): the caller's
    // 'b c (h w)' tensor read in place; resident-codebook tcgen05 kernel only
    int x_hw;
};

// Shared-memory carve-up, computed identically on host and device.
struct SmemPlan {
    int xt, et, e2s, xn2, sidx, amb, order, misc, red, hist, cntw, start, accs, total;
};
__host__ __device__ inline SmemPlan make_smem_plan(int dp, int k, int stats_mode, int use_hist, int code_rows) {
    SmemPlan p;
    int o = 0;
    p.xt = o;    o += kBM * dp * 4;
    p.et = o;    o += code_rows * dp * 4;
    p.e2s = o;   o += 256 * 4;
    p.xn2 = o;   o += kBM * 4;
    p.sidx = o;  o += kBM * 4;
    p.amb = o;   o += kBM * 4;
    p.order = o; o += kBM * 4;
    p.misc = o;  o += 16 * 4;
    p.red = o;   o += 16 * 8;
    p.hist = o;  o += use_hist ? ((k + 3) & ~3) * 4 : 0;
    p.cntw = o;  o += stats_mode == kStatsSmall ? 4 * ((k + 3) & ~3) * 4 : 0;
    p.start = o; o += stats_mode == kStatsSmall ? ((k + 1 + 3) & ~3) * 4 : 0;
    p.accs = o;  o += stats_mode == kStatsSmall ? k * dp * 4 : 0;
    p.total = o + 1024;  // slack for 1024-byte alignment of the dynamic base
    return p;
}

// ------------------------------------------------------------------------------------------
// Phase R: canonical fp64 scan of one row against codes [0, k) — warp-cooperative.
// x row comes from the shared tile, code words from global memory (L1/L2 resident).
template <int DP, int ROWS = kBM>
__device__ __forceinline__ int canon_scan_row(const float* xt, int row, const float* cb, const float* e2g, int k,
                                              int d, int lane) {
    const int nchunk = d >> 2;
    float4 xv[DP / 128 > 0 ? DP / 128 : 1];
    double p = 0.0;
#pragma unroll
    for (int i = 0; i < (DP / 128 > 0 ? DP / 128 : 1); ++i) {
        int c = lane + 32 * i;
        xv[i] = (c < nchunk) ? *reinterpret_cast<const float4*>(xt + tile_off<ROWS>(row, c)) : make_float4(0, 0, 0, 0);
        if (c < nchunk) p = dot4(p, xv[i], xv[i]);
    }
    const float x2 = __double2float_rn(butterfly_sum(p));
    float best = __int_as_float(0x7f800000);
    int arg = 0;
    for (int code = 0; code < k; ++code) {
        const float4* er = reinterpret_cast<const float4*>(cb + (size_t)code * d);
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < (DP / 128 > 0 ? DP / 128 : 1); ++i) {
            int c = lane + 32 * i;
            if (c < nchunk) s = dot4(s, xv[i], __ldg(er + c));
        }
        s = butterfly_sum(s);
        float dk = canon_score(x2, s, __ldg(e2g + code));
        if (dk < best) { best = dk; arg = code; }   // strict: first index wins ties (torch argmax rule)
    }
    return arg;
}

// Same, restricted to a short candidate list (used by the tcgen05 path).
template <int DP, int ROWS = kBM>
__device__ __forceinline__ int canon_pick(const float* xt, int row, const float* cb, const float* e2g, int d,
                                          int lane, const int* cand, int ncand) {
    const int nchunk = d >> 2;
    float4 xv[DP / 128 > 0 ? DP / 128 : 1];
    double p = 0.0;
#pragma unroll
    for (int i = 0; i < (DP / 128 > 0 ? DP / 128 : 1); ++i) {
        int c = lane + 32 * i;
        xv[i] = (c < nchunk) ? *reinterpret_cast<const float4*>(xt + tile_off<ROWS>(row, c)) : make_float4(0, 0, 0, 0);
        if (c < nchunk) p = dot4(p, xv[i], xv[i]);
    }
    const float x2 = __double2float_rn(butterfly_sum(p));
    float best = __int_as_float(0x7f800000);
    int arg = 0x7fffffff;
    for (int j = 0; j < ncand; ++j) {
        const int code = cand[j];
        const float4* er = reinterpret_cast<const float4*>(cb + (size_t)code * d);
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < (DP / 128 > 0 ? DP / 128 : 1); ++i) {
            int c = lane + 32 * i;
            if (c < nchunk) s = dot4(s, xv[i], __ldg(er + c));
        }
        s = butterfly_sum(s);
        float dk = canon_score(x2, s, __ldg(e2g + code));
        if (dk < best || (dk == best && code < arg)) { best = dk; arg = code; }
    }
    return arg;
}

// ------------------------------------------------------------------------------------------
// Phase A and the end-of-kernel flush are shared by both scoring paths.
struct ApplyState {
    float loss;  // per-thread partial of sum (q_st - x)^2
};

// Gather / straight-through / loss / large-k statistics for the rows of one tile.
template <int DP, bool TRAIN, int ROWS = kBM>
__device__ __forceinline__ void apply_rows(const FwdParams& p, const float* xt, const int* sidx, int* hist,
                                           int64_t row0, ApplyState& st, int warp = threadIdx.x >> 5,
                                           int nwarps = kWarps) {
    const int lane = threadIdx.x & 31;
    const int nchunk = p.d >> 2;
    float* esum = p.stats + ((p.k + 3) & ~3);
    for (int r = warp; r < ROWS; r += nwarps) {
        const int64_t grow = row0 + r;
        if (grow >= p.n) break;
        const int code = sidx[r];
        const float4* er = reinterpret_cast<const float4*>(p.cb + (size_t)code * p.d);
        for (int c = lane; c < nchunk; c += 32) {
            const float4 xv = *reinterpret_cast<const float4*>(xt + tile_off<ROWS>(r, c));
            const float4 ev = __ldg(er + c);
            float4 o;
            if (TRAIN) {
                // x + (e - x): two rounded fp32 ops, never contracted; the loss is
                // taken on that rounded tensor, as F.mse_loss(quantize.detach(), x) does.
                o.x = __fadd_rn(xv.x, __fsub_rn(ev.x, xv.x));
                o.y = __fadd_rn(xv.y, __fsub_rn(ev.y, xv.y));
                o.z = __fadd_rn(xv.z, __fsub_rn(ev.z, xv.z));
                o.w = __fadd_rn(xv.w, __fsub_rn(ev.w, xv.w));
                float dx = __fsub_rn(o.x, xv.x), dy = __fsub_rn(o.y, xv.y);
                float dz = __fsub_rn(o.z, xv.z), dw = __fsub_rn(o.w, xv.w);
                st.loss = fmaf(dx, dx, st.loss);
                st.loss = fmaf(dy, dy, st.loss);
                st.loss = fmaf(dz, dz, st.loss);
                st.loss = fmaf(dw, dw, st.loss);
                if (p.stats_mode == kStatsLarge) red_add_v4(esum + (size_t)code * p.d + 4 * c, xv);
            } else {
                o = ev;
            }
            if (p.q) st_stream_v4(p.q + (size_t)grow * p.d + 4 * c, o);
        }
        if (lane == 0 && p.stats_mode != kStatsSmall) {
            if (p.use_hist) atomicAdd(hist + code, 1);
            else atomicAdd(p.stats + code, 1.0f);
        }
    }
}

// Small-k statistics: rows of the tile are bucketed by code (deterministic counting sort), then
// every thread owns one column of a contiguous range of codes and adds its rows in bucket order
// into a private shared-memory slot — no floating-point atomics inside the CTA.
template <int DP>
__device__ __forceinline__ void small_stats_tile(const FwdParams& p, const float* xt, const int* sidx, int* hist,
                                                 int* cntw, int* start, int* order, float* accs, int64_t row0) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = p.k;
    const int kp = (k + 3) & ~3;
    for (int i = tid; i < 4 * kp; i += kThreads) cntw[i] = 0;
    __syncthreads();
    int code = -1, rank = 0;
    if (tid < kBM) {
        const bool valid = row0 + tid < p.n;
        code = valid ? sidx[tid] : -1;
        const unsigned m = __match_any_sync(0xffffffffu, code);
        rank = __popc(m & lanemask_lt());
        if (valid && rank == 0) cntw[warp * kp + code] = __popc(m);
    }
    __syncthreads();
    if (warp == 0) {   // exclusive scan of the per-code totals
        const int per = (k + 31) / 32;
        const int c0 = lane * per, c1 = min(k, c0 + per);
        int s = 0;
        for (int c = c0; c < c1; ++c) s += cntw[c] + cntw[kp + c] + cntw[2 * kp + c] + cntw[3 * kp + c];
        int incl = s;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            int v = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += v;
        }
        int run = incl - s;
        for (int c = c0; c < c1; ++c) {
            start[c] = run;
            run += cntw[c] + cntw[kp + c] + cntw[2 * kp + c] + cntw[3 * kp + c];
        }
        if (lane == 31) start[k] = incl;
    }
    __syncthreads();
    if (tid < kBM && code >= 0) {
        int pos = start[code] + rank;
        for (int w = 0; w < warp; ++w) pos += cntw[w * kp + code];
        order[pos] = tid;
    }
    __syncthreads();
    constexpr int G = kThreads / DP > 0 ? kThreads / DP : 1;   // column groups (DP <= 256)
    const int col = tid % DP, g = tid / DP;
    const int per = (k + G - 1) / G;
    const int c0 = g * per, c1 = min(k, c0 + per);
    const int xoff = (col & 3);
    const int xc4 = col >> 2;
    for (int c = c0; c < c1; ++c) {
        const int beg = start[c], end = start[c + 1];
        if (beg == end) continue;
        float a = 0.f;
        for (int i = beg; i < end; ++i) a += xt[tile_off<kBM>(order[i], xc4) + xoff];
        accs[c * DP + col] += a;
        if (col == 0) hist[c] += end - beg;
    }
    __syncthreads();
}

// Data-parallel statistics exchange, called by every thread of the LAST CTA of a launch (p.dp_world > 1): the one-shot
// all-reduce over NVLink peer memory of tvq_aux.cuh::ema_dp_kernel (same buffer layout and protocol: push the packed
// block into this rank's slot on every rank, fence.sys, publish flag = step everywhere, wait for all local flags, add the
// slots in rank order), with the rank-ordered sum written back into the local statistics p.stats — after it the
// single-rank update can run unchanged, and every rank holds identical bits.
__device__ __forceinline__ void dp_reduce_stats(const FwdParams& p, int* misc) {
    const int tid = threadIdx.x;
    const int world = p.dp_world;
    const int kp = (p.k + 3) & ~3;
    const int64_t len4 = (int64_t)(kp + p.k * p.d) >> 2;
    unsigned char* mine = reinterpret_cast<unsigned char*>(p.peers[p.dp_rank]);
    if (tid == 0) {
        unsigned* counter = reinterpret_cast<unsigned*>(mine);
        misc[2] = (int)(*counter + 1u);
        *counter = (unsigned)misc[2];
    }
    __syncthreads();
    const unsigned epoch = (unsigned)misc[2];
    const int par = (int)(epoch & 1u);
    const size_t flags_off = 64, slots_off = 64 + (((size_t)2 * world * 4 + 63) & ~(size_t)63);
    const float4* src = reinterpret_cast<const float4*>(p.stats);
    for (int r = 0; r < world; ++r) {
        float4* dst = reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(p.peers[r]) + slots_off) +
                      ((size_t)par * world + p.dp_rank) * len4;
        for (int64_t f = tid; f < len4; f += blockDim.x) dst[f] = __ldcg(src + f);
    }
    __threadfence_system();
    __syncthreads();
    if (tid < world) {
        unsigned* flag = reinterpret_cast<unsigned*>(reinterpret_cast<unsigned char*>(p.peers[tid]) + flags_off) + par * world + p.dp_rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
        const unsigned* lf = reinterpret_cast<const unsigned*>(mine + flags_off) + par * world + tid;
        wait_flag_sys(lf, epoch, p.dp_timeout_ns, reinterpret_cast<unsigned*>(mine) + 1);
    }
    __syncthreads();
    const float4* slots = reinterpret_cast<const float4*>(mine + slots_off) + (size_t)par * world * len4;
    float4* out = reinterpret_cast<float4*>(p.stats);
    for (int64_t f = tid; f < len4; f += blockDim.x) {
        float4 sv = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < world; ++r) {
            const float4 v = __ldcg(slots + (size_t)r * len4 + f);
            sv.x += v.x; sv.y += v.y; sv.z += v.z; sv.w += v.w;
        }
        out[f] = sv;
    }
    __syncthreads();
}

// The first half of dp_reduce_stats only: push this rank's packed statistics into its slot on every rank and publish the
// flags — no wait.  The matching second half (wait for all flags of the step, rank-ordered sum, EMA update) is
// tvq_aux.cuh::ema_dp_kernel in its finalize form, launched by the caller on a stream of its choice: the exchange latency
// and the inter-rank skew then overlap whatever the caller runs next (the backward kernels) instead of holding this
// kernel's last CTA.  Called by every thread of the LAST CTA.
__device__ __forceinline__ void dp_publish_stats(const FwdParams& p, int* misc) {
    const int tid = threadIdx.x;
    const int world = p.dp_world;
    const int kp = (p.k + 3) & ~3;
    const int64_t len4 = (int64_t)(kp + p.k * p.d) >> 2;
    unsigned char* mine = reinterpret_cast<unsigned char*>(p.peers[p.dp_rank]);
    if (tid == 0) {
        unsigned* counter = reinterpret_cast<unsigned*>(mine);
        misc[2] = (int)(*counter + 1u);
        *counter = (unsigned)misc[2];
    }
    __syncthreads();
    const unsigned epoch = (unsigned)misc[2];
    const int par = (int)(epoch & 1u);
    const size_t flags_off = 64, slots_off = 64 + (((size_t)2 * world * 4 + 63) & ~(size_t)63);
    const float4* src = reinterpret_cast<const float4*>(p.stats);
    for (int r = 0; r < world; ++r) {
        float4* dst = reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(p.peers[r]) + slots_off) +
                      ((size_t)par * world + p.dp_rank) * len4;
        for (int64_t f = tid; f < len4; f += blockDim.x) dst[f] = __ldcg(src + f);
    }
    __threadfence_system();
    __syncthreads();
    if (tid < world) {
        unsigned* flag = reinterpret_cast<unsigned*>(reinterpret_cast<unsigned char*>(p.peers[tid]) + flags_off) + par * world + p.dp_rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
    }
    __syncthreads();
}

// Last-CTA epilogue (threadFenceReduction pattern): the CTA that takes the final ticket turns the
// global totals into the scalars.  Called by every thread of every CTA after its flush.
template <bool TRAIN>
__device__ __forceinline__ void finish_ticket(const FwdParams& p, double* red, int* misc) {
    const int tid = threadIdx.x;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        unsigned t = atomicAdd(&p.hdr->ticket, 1u);
        misc[1] = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (misc[1]) {
        __threadfence();
        // perplexity = exp(-sum p log(p + 1e-10)), p = counts / n  (vq.py:246-247)
        const float fn = (float)p.n;
        float part = 0.f;
        for (int c = tid; c < p.k; c += blockDim.x) {
            float cnt = __ldcg(p.stats + c);
            float pr = __fdiv_rn(cnt, fn);
            part += pr * logf(pr + 1e-10f);
        }
        double tot = block_sum((double)part, red);
        if (tid == 0) {
            p.scalars[1] = expf(-(float)tot);
            double ls = *reinterpret_cast<volatile double*>(&p.hdr->loss_sum);
            const float commit = TRAIN ? (float)(ls / ((double)p.n * (double)p.d)) : 0.f;
            p.scalars[0] = commit;
            p.scalars[2] = __fmul_rn(commit, p.commitment_weight);
            p.scalars[3] = 0.f;
            reinterpret_cast<unsigned*>(p.scalars)[4] = *reinterpret_cast<volatile unsigned*>(&p.hdr->n_rescored);
            reinterpret_cast<unsigned*>(p.scalars)[5] = *reinterpret_cast<volatile unsigned*>(&p.hdr->n_exact);
            p.scalars[6] = 0.f;
            p.scalars[7] = 0.f;
            if (p.commit_out) *p.commit_out = commit;
            if (p.weighted_out) *p.weighted_out = __fmul_rn(commit, p.commitment_weight);
            p.hdr->ticket = 0;
            p.hdr->next_tile = 0u;
            p.hdr->loss_sum = 0.0;             // consumed: the header is clean for the next call
            p.hdr->n_rescored = 0u;
            p.hdr->n_exact = 0u;
        }
        if (TRAIN && p.fuse_ema) {
            // EMA codebook update (vq.py:231,236-242) by this last CTA: every other CTA has taken its ticket,
            // i.e. finished reading the codebook and flushing its statistics.  Same arithmetic as ema_kernel
            // (tvq_aux.cuh); the statistics scratch is zeroed for the next call.
            // Data-parallel (p.dp_world > 1): the statistics are first summed over the ranks by a one-shot
            // exchange over NVLink peer memory (protocol and buffer layout: tvq_aux.cuh::ema_dp_kernel).
            __syncthreads();
            const int kp = (p.k + 3) & ~3;
            const int dq = p.d >> 2;
            const int world = p.dp_world > 1 ? p.dp_world : 1;
            const int64_t len4 = (int64_t)(kp + p.k * p.d) >> 2;
            const float4* slots = reinterpret_cast<const float4*>(p.stats);     // world slots of len4 float4, rank order
            if (p.dp_world > 1) {
                unsigned char* mine = reinterpret_cast<unsigned char*>(p.peers[p.dp_rank]);
                if (tid == 0) {
                    unsigned* counter = reinterpret_cast<unsigned*>(mine);
                    misc[2] = (int)(*counter + 1u);
                    *counter = (unsigned)misc[2];
                }
                __syncthreads();
                const unsigned epoch = (unsigned)misc[2];
                const int par = (int)(epoch & 1u);
                const size_t flags_off = 64, slots_off = 64 + (((size_t)2 * world * 4 + 63) & ~(size_t)63);
                const float4* src = reinterpret_cast<const float4*>(p.stats);
                for (int r = 0; r < world; ++r) {
                    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(p.peers[r]) + slots_off) +
                                  ((size_t)par * world + p.dp_rank) * len4;
                    for (int64_t f = tid; f < len4; f += blockDim.x) dst[f] = __ldcg(src + f);
                }
                __threadfence_system();
                __syncthreads();
                if (tid < world) {
                    unsigned* flag = reinterpret_cast<unsigned*>(reinterpret_cast<unsigned char*>(p.peers[tid]) + flags_off) + par * world + p.dp_rank;
                    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
                    const unsigned* lf = reinterpret_cast<const unsigned*>(mine + flags_off) + par * world + tid;
                    wait_flag_sys(lf, epoch, p.dp_timeout_ns, reinterpret_cast<unsigned*>(mine) + 1);
                }
                __syncthreads();
                slots = reinterpret_cast<const float4*>(mine + slots_off) + (size_t)par * world * len4;
            }
            auto count_of = [&](int c) {
                float cnt = 0.f;
                for (int r = 0; r < world; ++r) cnt += __ldcg(reinterpret_cast<const float*>(slots + (size_t)r * len4) + c);
                return cnt;
            };
            double part = 0.0;
            for (int c = tid; c < p.k; c += blockDim.x)
                part += (double)fmaf(count_of(c), p.one_minus_decay, __fmul_rn(p.cluster_size[c], p.decay));
            const double totn = block_sum(part, red);
            if (tid == 0) red[15] = totn;
            __syncthreads();
            const float nsum = __double2float_rn(red[15]);
            const float denom = __fadd_rn(nsum, p.k_eps);
            float4* esum4 = reinterpret_cast<float4*>(p.stats + kp);
            float4* avg4 = reinterpret_cast<float4*>(p.embed_avg);
            float4* emb4 = reinterpret_cast<float4*>(p.embed);
            float4* prev4 = reinterpret_cast<float4*>(p.embed_prev);
            for (int f = tid; f < p.k * dq; f += blockDim.x) {
                const int c = f / dq;
                const float cs = fmaf(count_of(c), p.one_minus_decay, __fmul_rn(p.cluster_size[c], p.decay));
                const float sm = __fmul_rn(__fdiv_rn(__fadd_rn(cs, p.eps), denom), nsum);
                float4 sv = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int r = 0; r < world; ++r) {
                    const float4 v = __ldcg(slots + (size_t)r * len4 + (kp >> 2) + f);
                    sv.x += v.x; sv.y += v.y; sv.z += v.z; sv.w += v.w;
                }
                float4 a = avg4[f];
                a.x = fmaf(sv.x, p.one_minus_decay, __fmul_rn(a.x, p.decay));
                a.y = fmaf(sv.y, p.one_minus_decay, __fmul_rn(a.y, p.decay));
                a.z = fmaf(sv.z, p.one_minus_decay, __fmul_rn(a.z, p.decay));
                a.w = fmaf(sv.w, p.one_minus_decay, __fmul_rn(a.w, p.decay));
                avg4[f] = a;
                if (prev4) prev4[f] = emb4[f];
                emb4[f] = make_float4(__fdiv_rn(a.x, sm), __fdiv_rn(a.y, sm), __fdiv_rn(a.z, sm), __fdiv_rn(a.w, sm));
            }
            __syncthreads();
            for (int c = tid; c < kp; c += blockDim.x) {
                if (c < p.k) p.cluster_size[c] = fmaf(count_of(c), p.one_minus_decay, __fmul_rn(p.cluster_size[c], p.decay));
            }
            __syncthreads();                              // (the local statistics may be slot 0 of `slots`)
            for (int f = tid; f < p.k * dq; f += blockDim.x) esum4[f] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int c = tid; c < kp; c += blockDim.x) p.stats[c] = 0.f;
        }
    }
}

// End of kernel: flush CTA-local statistics and the loss partial; last CTA computes the scalars.
template <int DP, bool TRAIN>
__device__ __forceinline__ void flush_and_finish(const FwdParams& p, int* hist, const float* accs, double* red,
                                                 int* misc, ApplyState& st) {
    const int tid = threadIdx.x;
    __syncthreads();
    if (p.use_hist) {
        for (int c = tid; c < p.k; c += kThreads) {
            int v = hist[c];
            if (v) atomicAdd(p.stats + c, (float)v);
        }
    }
    if (TRAIN && p.stats_mode == kStatsSmall) {
        float* esum = p.stats + ((p.k + 3) & ~3);
        for (int f = tid; f < p.k * DP; f += kThreads) {
            const int c = f / DP, col = f % DP;
            const float v = accs[f];
            if (col < p.d && v != 0.f) atomicAdd(esum + (size_t)c * p.d + col, v);
        }
    }
    if (TRAIN) {
        double t = block_sum((double)st.loss, red);
        if (tid == 0) atomicAdd(&p.hdr->loss_sum, t);
    }
    finish_ticket<TRAIN>(p, red, misc);
}

// ------------------------------------------------------------------------------------------
template <int DP, bool TRAIN>
__global__ void __launch_bounds__(kThreads, (DP <= 128 ? 2 : 1)) fwd_simt_kernel(const FwdParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const SmemPlan pl = make_smem_plan(DP, p.k, p.stats_mode, p.use_hist, kBN);
    float* xt = reinterpret_cast<float*>(base + pl.xt);
    float* et = reinterpret_cast<float*>(base + pl.et);
    float* e2s = reinterpret_cast<float*>(base + pl.e2s);
    float* xn2 = reinterpret_cast<float*>(base + pl.xn2);
    int* sidx = reinterpret_cast<int*>(base + pl.sidx);
    int* amb = reinterpret_cast<int*>(base + pl.amb);
    int* order = reinterpret_cast<int*>(base + pl.order);
    int* misc = reinterpret_cast<int*>(base + pl.misc);
    double* red = reinterpret_cast<double*>(base + pl.red);
    int* hist = reinterpret_cast<int*>(base + pl.hist);
    int* cntw = reinterpret_cast<int*>(base + pl.cntw);
    int* start = reinterpret_cast<int*>(base + pl.start);
    float* accs = reinterpret_cast<float*>(base + pl.accs);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = tid & 7, ty = tid >> 3;
    constexpr int DPC = DP / 4;
    const int nchunk = p.d >> 2;
    const int ncode_tiles = (p.k + kBN - 1) / kBN;

    // ---- CTA prologue: zero CTA-local statistics, find max |e|^2 (error-bound constant)
    if (p.use_hist) for (int c = tid; c < p.k; c += kThreads) hist[c] = 0;
    if (TRAIN && p.stats_mode == kStatsSmall) for (int f = tid; f < p.k * DP; f += kThreads) accs[f] = 0.f;
    float emax2 = 0.f;
    for (int c = tid; c < p.k; c += kThreads) emax2 = fmaxf(emax2, __ldg(p.e2 + c));
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) emax2 = fmaxf(emax2, __shfl_xor_sync(0xffffffffu, emax2, off));
    if (lane == 0) xn2[warp] = emax2;
    __syncthreads();
    emax2 = 0.f;
    for (int w = 0; w < kWarps; ++w) emax2 = fmaxf(emax2, xn2[w]);
    const float emax = sqrtf(emax2) * 1.0001f;
    __syncthreads();
    // rigorous bound of |(s_a - s_b) - (d_a - d_b)| for fp32 FMA-chain scores
    const float err_c = (float)(DP + 16) * 5.9604645e-8f * 1.05f;

    ApplyState st;
    st.loss = 0.f;
    bool codes_resident = false;

    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int64_t row0 = (int64_t)tile * kBM;
        // ---------------- L: x tile -> shared memory
        for (int f = tid; f < kBM * DPC; f += kThreads) {
            const int row = f / DPC, c4 = f % DPC;
            float* dst = xt + tile_off<kBM>(row, c4);
            const int64_t grow = row0 + row;
            if (grow < p.n && c4 < nchunk) cp_async16(dst, p.x + (size_t)grow * p.d + 4 * c4);
            else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        cp_async_commit();
        if (tid == 0) misc[0] = 0;   // ambiguous-row count of this tile
        cp_async_wait_all();
        __syncthreads();
        // row norms (for the error bound only; any summation order will do)
        for (int r = warp; r < kBM; r += kWarps) {
            float s = 0.f;
            for (int c = lane; c < DPC; c += 32) {
                float4 v = *reinterpret_cast<const float4*>(xt + tile_off<kBM>(r, c));
                s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            if (lane == 0) xn2[r] = s;
        }
        // ---------------- S: fp32 scoring, running top-2 per row
        const float INF = __int_as_float(0x7f800000);
        float rm1[4], rm2[4];
        int ri1[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { rm1[i] = INF; rm2[i] = INF; ri1[i] = 0; }
        if (!p.exact && !p.given_idx) {
            for (int ct = 0; ct < ncode_tiles; ++ct) {
                if (!codes_resident) {
                    __syncthreads();   // previous users of et / e2s are done
                    for (int f = tid; f < kBN * DPC; f += kThreads) {
                        const int row = f / DPC, c4 = f % DPC;
                        const int code = ct * kBN + row;
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (code < p.k && c4 < nchunk) v = __ldg(reinterpret_cast<const float4*>(p.cb + (size_t)code * p.d) + c4);
                        *reinterpret_cast<float4*>(et + tile_off<kBN>(row, c4)) = v;
                    }
                    if (tid < kBN) e2s[tid] = (ct * kBN + tid < p.k) ? __ldg(p.e2 + ct * kBN + tid) : INF;
                    __syncthreads();
                    codes_resident = (ncode_tiles == 1);
                }
                float acc[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 2
                for (int c4 = 0; c4 < DPC; ++c4) {
                    float4 xa[4], eb[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) xa[i] = *reinterpret_cast<const float4*>(xt + tile_off<kBM>(ty + 32 * i, c4));
#pragma unroll
                    for (int j = 0; j < 4; ++j) eb[j] = *reinterpret_cast<const float4*>(et + tile_off<kBN>(tx + 8 * j, c4));
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            acc[i][j] = fmaf(xa[i].x, eb[j].x, acc[i][j]);
                            acc[i][j] = fmaf(xa[i].y, eb[j].y, acc[i][j]);
                            acc[i][j] = fmaf(xa[i].z, eb[j].z, acc[i][j]);
                            acc[i][j] = fmaf(xa[i].w, eb[j].w, acc[i][j]);
                        }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float m1 = INF, m2 = INF;
                    int i1 = 0x7fffffff;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float s = fmaf(-2.f, acc[i][j], e2s[tx + 8 * j]);
                        const int code = ct * kBN + tx + 8 * j;
                        if (s < m1) { m2 = m1; m1 = s; i1 = code; }
                        else if (s < m2) { m2 = s; }
                    }
#pragma unroll
                    for (int off = 1; off <= 4; off <<= 1) {
                        const float om1 = __shfl_xor_sync(0xffffffffu, m1, off);
                        const float om2 = __shfl_xor_sync(0xffffffffu, m2, off);
                        const int oi1 = __shfl_xor_sync(0xffffffffu, i1, off);
                        if (om1 < m1 || (om1 == m1 && oi1 < i1)) { m2 = fminf(m1, om2); m1 = om1; i1 = oi1; }
                        else { m2 = fminf(m2, om1); }
                    }
                    if (m1 < rm1[i]) { rm2[i] = fminf(rm1[i], m2); rm1[i] = m1; ri1[i] = i1; }
                    else { rm2[i] = fminf(rm2[i], m1); }
                }
            }
        }
        // ---------------- R: classify rows; canonical scan for the undecidable ones
        __syncthreads();   // row norms visible
        if (p.given_idx) {
            if (tid < kBM) {
                int64_t v = (row0 + tid < p.n) ? p.idx[row0 + tid] : 0;
                sidx[tid] = (int)(v < 0 ? 0 : (v >= p.k ? p.k - 1 : v));
            }
        } else if (tx == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = ty + 32 * i;
                const float b = sqrtf(xn2[row]) * 1.0001f + emax;
                const float thr = err_c * b * b;
                const bool decided = (rm2[i] - rm1[i] > thr) && !p.exact;   // NaN -> not decided
                sidx[row] = ri1[i];
                if (!decided && row0 + row < p.n) amb[atomicAdd(&misc[0], 1)] = row;
            }
        }
        __syncthreads();
        const int namb = misc[0];
        for (int a = warp; a < namb; a += kWarps) {
            const int row = amb[a];
            const int best = canon_scan_row<DP>(xt, row, p.cb, p.e2, p.k, p.d, lane);
            if (lane == 0) sidx[row] = best;
        }
        if (tid == 0 && namb) { atomicAdd(&p.hdr->n_rescored, (unsigned)namb); atomicAdd(&p.hdr->n_exact, (unsigned)namb); }
        __syncthreads();
        // ---------------- A: outputs and statistics
        if (!p.given_idx && tid < kBM && row0 + tid < p.n) p.idx[row0 + tid] = (int64_t)sidx[tid];
        apply_rows<DP, TRAIN>(p, xt, sidx, hist, row0, st);
        if (TRAIN && p.stats_mode == kStatsSmall) small_stats_tile<DP>(p, xt, sidx, hist, cntw, start, order, accs, row0);
        __syncthreads();   // xt is free for the next tile
    }
    flush_and_finish<DP, TRAIN>(p, hist, accs, red, misc, st);
}

}  // namespace tvq
