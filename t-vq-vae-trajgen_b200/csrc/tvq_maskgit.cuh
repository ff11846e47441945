// tvq_maskgit.cuh — ONE MaskGIT decoding iteration after the transformer, in ONE kernel:
// /root/reference/timevqvae/models/maskgit.py:300-346 (first_pass loop body; second_pass :364-410 is the same) with
// mask_by_random_topk :238-267.  Per sequence: sample a token id for every position from softmax(logits) —
// Categorical.sample() is torch.multinomial with one draw, i.e. argmax(probs / q), q ~ Exp(1) —, keep the
// already-decoded tokens, take the probability of the sampled id as confidence (inf for known tokens), add
// temperature * Gumbel noise to its log, and re-mask the `mask_len` least confident positions.  The reference runs a
// dozen elementwise / reduction kernels, a top-k and a Python loop over the batch for this; here one CTA owns one
// sequence.  The noise (q, u) is an INPUT: the host draws it from torch's generator in the reference's order, so the
// result is the reference's for the same generator state (up to positions whose two best ratios / confidences differ
// by less than the rounding of expf / logf, which no two float implementations agree on).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tvq {

struct MaskgitParams {
    const float* logits;     // [b, n, k]
    const int64_t* s;        // [b, n] current tokens (mask_token_id = still unknown)
    const float* q;          // [b, n, k] Exp(1) noise of the categorical draw
    const float* u;          // [b, n]    U(0,1) noise of the Gumbel perturbation
    int64_t b;
    int n, k;
    int64_t mask_token_id;
    int mask_len;
    float temperature;
    int64_t* s_new;          // [b, n]
    int64_t* sampled;        // [b, n] or null: ids before re-masking
    uint8_t* masking;        // [b, n] or null
};

__global__ void __launch_bounds__(128) maskgit_step_kernel(const MaskgitParams p) {
    extern __shared__ float msm[];
    float* conf = msm;                                   // [n]
    int* samp = reinterpret_cast<int*>(msm + p.n);       // [n]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const float INF = __int_as_float(0x7f800000);
    for (int64_t row = blockIdx.x; row < p.b; row += gridDim.x) {
        __syncthreads();
        // ---- per position: softmax, categorical draw (argmax p / q), confidence
        for (int i = warp; i < p.n; i += nwarps) {
            const float* lg = p.logits + (row * p.n + i) * (int64_t)p.k;
            const float* qq = p.q + (row * p.n + i) * (int64_t)p.k;
            float m = -INF;
            for (int j = lane; j < p.k; j += 32) m = fmaxf(m, __ldg(lg + j));
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
            float z = 0.f;
            for (int j = lane; j < p.k; j += 32) z += expf(__ldg(lg + j) - m);
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) z += __shfl_xor_sync(0xffffffffu, z, off);
            float best = -INF, bestp = 0.f;
            int arg = 0x7fffffff;
            for (int j = lane; j < p.k; j += 32) {
                const float pj = __fdiv_rn(expf(__ldg(lg + j) - m), z);
                const float r = __fdiv_rn(pj, __ldg(qq + j));
                if (r > best) { best = r; arg = j; bestp = pj; }          // strict: first index wins within the lane
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, off);
                const float op = __shfl_xor_sync(0xffffffffu, bestp, off);
                const int oa = __shfl_xor_sync(0xffffffffu, arg, off);
                if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; bestp = op; }
            }
            if (lane == 0) {
                const int64_t cur = __ldg(p.s + row * p.n + i);
                const bool unknown = cur == p.mask_token_id;
                const float sel = unknown ? bestp : INF;
                const float uu = __ldg(p.u + row * p.n + i);
                const float gumbel = -logf(fmaxf(-logf(fmaxf(uu, 1e-20f)), 1e-20f));
                conf[i] = logf(sel + 1e-5f) + p.temperature * gumbel;
                samp[i] = unknown ? (arg == 0x7fffffff ? 0 : arg) : (int)cur;
            }
        }
        __syncthreads();
        // ---- re-mask the mask_len least confident positions (rank by confidence, lower index first on ties)
        for (int i = tid; i < p.n; i += blockDim.x) {
            const float ci = conf[i];
            int rank = 0;
            for (int j = 0; j < p.n; ++j) {
                const float cj = conf[j];
                rank += (cj < ci || (cj == ci && j < i)) ? 1 : 0;
            }
            const bool masked = rank < p.mask_len;
            const int64_t o = row * p.n + i;
            p.s_new[o] = masked ? p.mask_token_id : (int64_t)samp[i];
            if (p.sampled) p.sampled[o] = (int64_t)samp[i];
            if (p.masking) p.masking[o] = masked ? 1 : 0;
        }
    }
}

}  // namespace tvq
