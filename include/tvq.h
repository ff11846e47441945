/* tvq.h — C ABI of the B200 (sm_100a) vector-quantisation kernels.
 *
 * This is the drop-in boundary underneath the reference's Python seam,
 * `timevqvae.models.vq.VectorQuantize` (/root/reference/timevqvae/models/vq.py:255-407):
 * the host module in t-vq-vae-trajgen_b200/vq.py keeps that class surface and calls these
 * entry points through ctypes.  Plain pointers and sizes only — no torch types, no C++
 * exceptions; every function returns 0 (TVQ_OK) or a negative tvq error / positive
 * cudaError_t.  All pointers are DEVICE pointers on the current CUDA device unless
 * marked host.  Kernels are enqueued on `stream` (a cudaStream_t passed as void*) and
 * never allocate, free or synchronise; buffers are owned by the caller.
 *
 * Shapes:  x, q      [n, d] fp32 row-major (the reference's `flatten`, vq.py:200)
 *          codebook  [k, d] fp32 row-major (`_codebook.embed`, vq.py:165)
 *          idx       [n]    int64          (`embed_ind`, vq.py:218-224)
 *          stats     [TVQ_STATS_OFFSET(k) + k*d] fp32: counts[k] (zero-padded to a multiple of
 *                    4 so that embed_sum stays 16-byte aligned) then embed_sum[k][d] — the two
 *                    tensors the reference all-reduces at vq.py:229 and :234 (embed_sum stored
 *                    K-major, i.e. already transposed as vq.py:236 uses it), packed so that
 *                    data-parallel ranks need ONE all-reduce
 */
#ifndef TVQ_H_
#define TVQ_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define TVQ_API __attribute__((visibility("default")))
#else
#define TVQ_API
#endif

#define TVQ_OK 0
#define TVQ_ERR_UNSUPPORTED (-1) /* shape outside the kernels' range (d % 4 != 0, d > 256, k < 1 ...) */
#define TVQ_ERR_BAD_ARG (-2)     /* null pointer, misaligned pointer, workspace too small */
#define TVQ_ERR_DEVICE (-3)      /* not an sm_100 device */

/* element offset of embed_sum inside `stats`, and the total element count */
#define TVQ_STATS_OFFSET(k) ((((int64_t)(k)) + 3) & ~(int64_t)3)
#define TVQ_STATS_LEN(k, d) (TVQ_STATS_OFFSET(k) + (int64_t)(k) * (int64_t)(d))

#define TVQ_NUM_SCALARS 8

/* flags for tvq_forward */
#define TVQ_F_TRAIN 1u      /* straight-through output + commitment-loss sum + embed_sum statistics  */
#define TVQ_F_WRITE_Q 2u    /* write q (eval: codebook[idx]; train: x + (codebook[idx] - x))         */
#define TVQ_F_EXACT 4u      /* debug: decide every row with the fp64 canonical scan (no fast scoring) */
#define TVQ_F_NO_UMMA 8u    /* force the SIMT scoring path even where the tcgen05 path applies        */
#define TVQ_F_GIVEN_IDX 16u /* idx is an INPUT (codes sampled by the caller, vq.py:55-56): skip scoring */

/* Library / device probes. */
TVQ_API int tvq_abi_version(void);
TVQ_API const char *tvq_error_string(int code);
/* 0 if `device` is compute capability 10.x; fills sm_count (may be NULL). */
TVQ_API int tvq_device_check(int device, int *sm_count);

/* One-shot launch hint of the calling thread: the next resident-codebook forward (tvq_forward / tvq_train_step* with
 * k <= 32 train / 64 eval, d <= 128) launched from this thread uses at most max_ctas CTAs; cleared by that launch (0 clears
 * it).  Results do not depend on the grid size (persistent kernel, dynamic tile scheduler).  For callers that run two
 * independent quantisers on two streams: with the SMs shared out (e.g. in proportion to the latents) the two launches are
 * resident together instead of one after the other.                                                              */
TVQ_API int tvq_hint_max_ctas(int max_ctas);

/* Bytes of scratch for (n, k, d): header + per-code constants (|e|^2 padded to a multiple of 256)
 * + tvq_train_step's statistics + the bf16 copy of the codebook that the streamed-codebook tcgen05
 * path feeds to TMA (k x roundup(d, 64..256) x 2 bytes, rewritten by every call).
 * The scratch must be zero-filled ONCE when allocated (it holds a launch ticket). */
TVQ_API size_t tvq_workspace_bytes(int64_t n, int k, int d);

/* Distance + assignment (+ gather / straight-through / loss / EMA statistics).
 *   Kernel selection: k <= 32 (train) / <= 64 (eval) and d <= 128: codebook resident in shared memory,
 *   tcgen05 tf32 scoring; every other shape: codebook streamed by TMA, tcgen05 bf16 scoring; the
 *   TVQ_F_EXACT / TVQ_F_NO_UMMA / TVQ_F_GIVEN_IDX flags select the CUDA-core kernel.  All three decide
 *   by the same canonical fp32/fp64 rule, so the indices do not depend on the path.
 *   Replaces EuclideanCodebook.forward vq.py:210-234 and the ST + commit-loss lines
 *   VectorQuantize.forward vq.py:357-366, for temperature 0.
 *   idx      out [n]
 *   q        out [n,d] or NULL (needs TVQ_F_WRITE_Q)
 *   stats    out [TVQ_STATS_LEN(k,d)]; zeroed inside; counts always, embed_sum only with TVQ_F_TRAIN
 *   scalars  out float[TVQ_NUM_SCALARS]: [0] commit loss = mean((q_st - x)^2) (train),
 *            [1] perplexity (vq.py:246-247, from this call's counts), [2] commitment_weight *
 *            commit loss (the reference's vq_loss["loss"], vq.py:366), [3] unused, [4],[5] uint32
 *            diagnostics: rows decided by the fp64 re-score, rows that needed the full exact scan. */
TVQ_API int tvq_forward(const float *x, const float *codebook, int64_t n, int k, int d, unsigned flags,
                float commitment_weight, int64_t *idx, float *q, float *stats, float *scalars,
                void *workspace, size_t workspace_bytes, void *stream);

/* One training-mode codebook step in as few launches as the shape allows: everything tvq_forward
 * does with TVQ_F_TRAIN | TVQ_F_WRITE_Q, then the EMA update of vq.py:227-242 on the statistics of
 * THIS call (no all-reduce in between: the single-GPU / sync_codebook=False path).  For k <= 32,
 * d <= 128 it is ONE kernel: the last CTA to finish applies the update.
 *   embed [k,d] is read as the codebook and updated in place; cluster_size [k], embed_avg [k,d]
 *   updated in place; embed_prev [k,d] (optional) receives the pre-update codebook for
 *   tvq_backward; commit_out / weighted_out (optional, 1 float each) receive scalars[0] / [2].
 *   The statistics live in a private part of `workspace` (zero before and after the call).    */
TVQ_API int tvq_train_step(const float *x, float *embed, float *cluster_size, float *embed_avg,
                   float *embed_prev, int64_t n, int k, int d, float commitment_weight,
                   double decay, double eps, int64_t *idx, float *q, float *scalars,
                   float *commit_out, float *weighted_out, void *workspace,
                   size_t workspace_bytes, void *stream);

/* EMA codebook update from (all-reduced) statistics: vq.py:231, :236-242 with helpers :59-64.
 *   cluster_size [k], embed_avg [k,d], embed [k,d] are updated in place; embed_prev [k,d]
 *   (optional, may be NULL) receives the codebook as it was before the update, which is the
 *   one the forward's outputs were gathered from (vq.py:225 precedes :242) and the one
 *   tvq_backward needs.                                                                      */
TVQ_API int tvq_ema_update(const float *stats, float *cluster_size, float *embed_avg, float *embed,
                   float *embed_prev, int k, int d, double decay, double eps, void *workspace,
                   size_t workspace_bytes, void *stream);

/* Data-parallel EMA update in ONE kernel (no NCCL call): one-shot all-reduce of the packed statistics
 * over NVLink peer memory fused with the EMA update — the reference's two all_reduce hooks
 * (vq.py:229, :234) plus vq.py:231, :236-242.
 *   peer_bufs  DEVICE array of `world` pointers: peer_bufs[r] is rank r's exchange buffer of
 *              tvq_exchange_bytes(k, d, world) bytes, allocated symmetrically, mapped by every rank
 *              (e.g. torch.distributed._symmetric_memory) and zero-filled ONCE before the first call.
 *   Every rank must call it the same number of times (the step counter lives in the buffer, so the
 *   launch can be replayed from a CUDA graph).  Slots are added in rank order on every rank, so the
 *   replicas stay bit-identical.  Statistics of at most 65 536 floats (TVQ_ERR_UNSUPPORTED beyond:
 *   use an all-reduce + tvq_ema_update).  stats (this rank's, from tvq_forward) must be padded to a
 *   multiple of 4 floats (TVQ_STATS_LEN rounded up).                                              */
TVQ_API size_t tvq_exchange_bytes(int k, int d, int world);
/* Exchange-buffer header (first 64 bytes of every rank's buffer, local use only): u32 [0] step counter, u32 [1] error word.
 * A data-parallel kernel waits for every peer's statistics at most `seconds` (default 1800; 0 = for ever).  When a peer does
 * not show up in time the kernel stores the step number in the error word, gives up the wait and finishes the step with
 * whatever that peer's slot holds: the CUDA context survives and the HOST decides (read the word; 0 = no error so far). */
TVQ_API int tvq_set_peer_timeout(double seconds);
TVQ_API int tvq_ema_update_dp(const float *stats, void *const *peer_bufs, int rank, int world,
                      float *cluster_size, float *embed_avg, float *embed, float *embed_prev, int k,
                      int d, double decay, double eps, void *stream);

/* tvq_train_step for data-parallel ranks, still ONE kernel per codebook (k <= 32, d <= 128; otherwise
 * TVQ_ERR_UNSUPPORTED: use tvq_forward + tvq_ema_update_dp): the last CTA to finish sums the statistics
 * of all ranks through the exchange buffers (same protocol and buffers as tvq_ema_update_dp) and applies
 * the EMA update.  Every rank must call it, with n >= 1.                                            */
TVQ_API int tvq_train_step_dp(const float *x, float *embed, float *cluster_size, float *embed_avg,
                      float *embed_prev, int64_t n, int k, int d, float commitment_weight, double decay,
                      double eps, int64_t *idx, float *q, float *scalars, float *commit_out,
                      float *weighted_out, void *workspace, size_t workspace_bytes,
                      void *const *peer_bufs, int rank, int world, void *stream);

/* Deferred exchange: the data-parallel step split in two so that the exchange latency and the skew between the ranks
 * are off the critical path.  tvq_hint_defer_exchange(mode) is a one-shot hint of the calling thread for the NEXT fused
 * data-parallel train step (tvq_train_step_dp / _qcf / _cf with world > 1) launched from this thread: the step fills idx, q,
 * scalars and embed_prev as usual but neither waits for the peers nor updates cluster_size / embed_avg / embed;
 *   mode 1: its last CTA still PUBLISHES this rank's statistics to the peers (remote stores + flags);
 *   mode 2: it leaves them in the workspace scratch — its tail is then exactly the single-GPU one.
 * tvq_ema_finalize_dp(published = (mode == 1), ...) completes the step: (mode 2: publish,) wait for every rank's
 * statistics of that step, add them in rank order, apply the EMA update.  It must be enqueued after the step it completes,
 * on any stream ordered after it, and before the next step on the same workspace / exchange buffers; every rank makes the
 * same sequence of calls.  The caller orders later readers of the three buffers after it.                       */
TVQ_API int tvq_hint_defer_exchange(int mode);
TVQ_API int tvq_ema_finalize_dp(int published, void *workspace, size_t workspace_bytes, void *const *peer_bufs, int rank,
                        int world, float *cluster_size, float *embed_avg, float *embed, int k, int d, double decay,
                        double eps, void *stream);

/* tvq_forward / tvq_train_step(_dp) writing q CHANNELS-FIRST: q [n / q_hw, d, q_hw] — the 'b c (h w)' layout of the
 * caller (utils/train_utils.py:349), so that quantize() needs no transpose after the VQ.  x stays [n, d]; idx stays
 * [n].  Resident-codebook kernel only (k <= 32 train / 64 eval, d <= 128), n % q_hw == 0; TVQ_ERR_UNSUPPORTED
 * otherwise (use the row-major call + tvq_transpose).  world = 1: local EMA update (peer_bufs may be NULL).       */
TVQ_API int tvq_forward_qcf(const float *x, const float *codebook, int64_t n, int k, int d, unsigned flags,
                    float commitment_weight, int64_t *idx, float *q, float *stats, float *scalars,
                    void *workspace, size_t workspace_bytes, int q_hw, void *stream);
TVQ_API int tvq_train_step_qcf(const float *x, float *embed, float *cluster_size, float *embed_avg,
                       float *embed_prev, int64_t n, int k, int d, float commitment_weight, double decay,
                       double eps, int64_t *idx, float *q, float *scalars, float *commit_out,
                       float *weighted_out, void *workspace, size_t workspace_bytes,
                       void *const *peer_bufs, int rank, int world, int q_hw, void *stream);
/* Backward for the same caller: g_zq and g_z are channels-first [b, d, hw], x [b*hw, d] and idx [b*hw] row-major
 * (as the forward saw them).  One kernel instead of transpose + tvq_backward + transpose.                        */
TVQ_API int tvq_backward_cf(const float *g_zq, const float *g_commit, const float *g_weighted, const float *x,
                    const int64_t *idx, const float *codebook, int64_t b, int hw, int k, int d,
                    float commitment_weight, float *g_z, void *stream);

/* The call site's own layout on BOTH sides: z and q are the encoder output / decoder input
 * 'b c (h w)' tensors [b, d, hw] of utils/train_utils.py:338-358, read and written in place — neither of the two
 * rearrange copies of :347,:349 exists.  idx is [b * hw] (latent b*hw + t).  Resident-codebook kernel only
 * (k <= 32 train / 64 eval, d <= 128, b * hw < 2^31 - 64); TVQ_ERR_UNSUPPORTED otherwise.  tvq_forward_cf: q may be
 * NULL iff TVQ_F_WRITE_Q is clear (tokenise).  tvq_backward_cfx is the matching backward (z, g_zq, g_z all [b, d, hw]). */
TVQ_API int tvq_forward_cf(const float *z, const float *codebook, int64_t b, int hw, int k, int d, unsigned flags,
                   float commitment_weight, int64_t *idx, float *q, float *stats, float *scalars,
                   void *workspace, size_t workspace_bytes, void *stream);
TVQ_API int tvq_train_step_cf(const float *z, float *embed, float *cluster_size, float *embed_avg,
                      float *embed_prev, int64_t b, int hw, int k, int d, float commitment_weight, double decay,
                      double eps, int64_t *idx, float *q, float *scalars, float *commit_out,
                      float *weighted_out, void *workspace, size_t workspace_bytes,
                      void *const *peer_bufs, int rank, int world, void *stream);
TVQ_API int tvq_backward_cfx(const float *g_zq, const float *g_commit, const float *g_weighted, const float *z,
                     const int64_t *idx, const float *codebook, int64_t b, int hw, int k, int d,
                     float commitment_weight, float *g_z, void *stream);

/* Backward of the train forward (autograd through vq.py:357-366):
 *   g_x = g_q + (g_commit + commitment_weight * g_weighted) * 2/(n*d) * (x - q_st)
 *   with q_st recomputed from x, idx and the codebook the forward used.  g_commit / g_weighted
 *   are device scalars: the gradients w.r.t. the commit loss (scalars[0]) and the weighted loss
 *   (scalars[2]); g_q is the gradient w.r.t. q.  Any of the three may be NULL (= zero).      */
TVQ_API int tvq_backward(const float *g_q, const float *g_commit, const float *g_weighted,
                 const float *x, const int64_t *idx,
                 const float *codebook, int64_t n, int k, int d, float commitment_weight,
                 float *g_x, void *stream);

/* Codeword gather for de-tokenising, F.embedding + rearrange of
 * /root/reference/timevqvae/models/maskgit.py:465-470.
 *   tokens [b*t] int64; layout 0: out[b, t, d]; layout 1: out[b, d, t] (decoder layout).      */
TVQ_API int tvq_gather(const int64_t *tokens, const float *codebook, int64_t b, int64_t t, int k, int d,
               int layout, float *out, void *stream);
/* The same, reporting ids outside [0, k) — F.embedding raises on them (e.g. a mask token, id == k, left over by an
 * incomplete MaskGIT pass): such a token's output row is NaN and *bad_count (device uint32, caller-zeroed, may be NULL)
 * is incremented once per offending token.  tvq_gather is this call with bad_count = NULL.                      */
TVQ_API int tvq_gather_checked(const int64_t *tokens, const float *codebook, int64_t b, int64_t t, int k, int d,
                       int layout, float *out, unsigned *bad_count, void *stream);

/* Full negative squared-distance matrix dist[n,k] (vq.py:210-214) for the stochastic
 * `svq_temp` branch (vq.py:55-56), whose sampling stays in torch to share its RNG stream.   */
TVQ_API int tvq_neg_dist(const float *x, const float *codebook, int64_t n, int k, int d, float *dist,
                 void *stream);

/* Stage-1 STFT front end in ONE kernel: everything the reference derives from a batch of
 * trajectories x [b, c, l] before the encoders run.  Any output may be NULL.
 *   xf        [b, 2c, n_fft/2+1, l/hop+1]  time_to_timefreq(x, n_fft, c)          utils/train_utils.py:293-307
 *   enc_in_l  (same shape)                 zero_pad_high_freq(xf, copy=True)      :361-372 (models/vq_vae.py:179-180)
 *   enc_in_h  (same shape)                 zero_pad_low_freq(xf, copy=True)       :375-386
 *   x_l, x_h  [b, c, l]                    F.interpolate(timefreq_to_time(zero_pad_{high,low}_freq(xf)), l, "linear")
 *                                          trainers/stage1.py:101-113
 * hop = n_fft / 4, periodic Hann window, reflect-centred, onesided, normalized — the arguments the reference
 * passes to torch.stft / torch.istft.  n_fft a multiple of 4 in [4, 64], l > n_fft / 2.                      */
TVQ_API int tvq_frontend(const float *x, int64_t b, int c, int l, int n_fft, float *xf, float *enc_in_l,
                 float *enc_in_h, float *x_l, float *x_h, void *stream);

/* Decoder side of the STFT front end (models/vq_vae.py:259-262, utils/train_utils.py:310-321, :361-386):
 *   y = F.interpolate(timefreq_to_time(pad_func(u), n_fft, c), l, "linear")   for u [b, 2c, n_fft/2+1, l/hop+1]
 * with pad_func = identity (band 0), zero_pad_high_freq (band 1: keep bin 0) or zero_pad_low_freq (band 2: keep
 * bins 1..), in ONE kernel; the backward (adjoint) writes g_u, zero in the bands pad_func removed.            */
TVQ_API int tvq_band_istft(const float *u, int64_t b, int c, int l, int n_fft, int band, float *y, void *stream);
/* The general form: u [b, 2c, n_fft/2+1, t] with ANY number of frames t >= 2 (the shipped decoders emit t = 384 (LF) / 400
 * (HF) frames for l = 200): the ISTFT yields hop * (t - 1) samples, F.interpolate(mode="linear", align_corners=False) maps
 * them onto l.  tvq_band_istft(_backward) is this call with t = l / hop + 1.                                    */
TVQ_API int tvq_band_istft_frames(const float *u, int64_t b, int c, int t, int l, int n_fft, int band, float *y,
                          void *stream);
TVQ_API int tvq_band_istft_frames_backward(const float *g_y, int64_t b, int c, int t, int l, int n_fft, int band,
                                   float *g_u, void *stream);
TVQ_API int tvq_band_istft_backward(const float *g_y, int64_t b, int c, int l, int n_fft, int band, float *g_u,
                            void *stream);

/* ONE MaskGIT decoding iteration after the transformer, one kernel:
 * /root/reference/timevqvae/models/maskgit.py:300-346 (= :364-410) with mask_by_random_topk :238-267.
 *   logits [b,n,k]; s [b,n] current tokens (mask_token_id = unknown); q [b,n,k] Exp(1) noise of the categorical draw
 *   (Categorical.sample() = argmax(probs / q)); u [b,n] U(0,1) noise of the Gumbel perturbation; mask_len positions
 *   with the lowest log(p + 1e-5) + temperature * Gumbel(u) are re-masked (known tokens have confidence inf).
 *   s_new [b,n] out; sampled [b,n] (ids before re-masking) and masking [b,n] (uint8) optional.               */
TVQ_API int tvq_maskgit_step(const float *logits, const int64_t *s, const float *q, const float *u, int64_t b, int n,
                     int k, int64_t mask_token_id, int mask_len, float temperature, int64_t *s_new,
                     int64_t *sampled, uint8_t *masking, void *stream);

/* Batched 2-D transpose in [b, r, s] -> out [b, s, r]: the layout change around the VQ in quantize(),
 * utils/train_utils.py:346-349 ('b c h w -> b (h w) c' and back), as a coalesced tiled copy.                  */
TVQ_API int tvq_transpose(const float *in, int64_t b, int r, int s, float *out, void *stream);

/* Snake activation of the stage-1 conv stacks, y = x + sin^2(a_c x) / a_c (utils/train_utils.py:421-448), forward and
 * backward in one kernel each — for the stage-1 harness, not the VQ hot path.  x, y, g, g_x [n, c, s] fp32 in NCHW order
 * (channels_last = 0) or NHWC order (1); a, g_a [c]; g_a is ACCUMULATED into (zero it first).  c <= 1024.          */
TVQ_API int tvq_snake_forward(const float *x, const float *a, int64_t n, int c, int64_t s, int channels_last, float *y,
                      void *stream);
TVQ_API int tvq_snake_backward(const float *g, const float *x, const float *a, int64_t n, int c, int64_t s,
                       int channels_last, float *g_x, float *g_a, void *stream);

/* Dead-code re-seed (vq.py:181-195): embed[j] = x[rows[j]] where cluster_size[j] < threshold.
 * Only `embed` is touched, as in the reference.  rows [k] int64 (drawn by the host).          */
TVQ_API int tvq_reseed(const float *x, const int64_t *rows, const float *cluster_size, float threshold,
               float *embed, int64_t n, int k, int d, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TVQ_H_ */
