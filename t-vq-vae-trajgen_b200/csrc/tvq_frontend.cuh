// tvq_frontend.cuh — stage-1 STFT front end in ONE kernel: everything the reference derives
// from a batch of trajectories before the encoders run,
//   xf        = time_to_timefreq(x)                      utils/train_utils.py:293-307  (torch.stft, hop = n_fft/4,
//                                                        periodic Hann, reflect-centred, onesided, normalized)
//   enc_in_l  = zero_pad_high_freq(xf, copy=True)        :361-372   (LF encoder input, models/vq_vae.py:179-180)
//   enc_in_h  = zero_pad_low_freq(xf, copy=True)         :375-386   (HF encoder input)
//   x_l       = interpolate(timefreq_to_time(zero_pad_high_freq(xf)), L)   trainers/stage1.py:101-107 (LF target)
//   x_h       = interpolate(timefreq_to_time(zero_pad_low_freq(xf)), L)    trainers/stage1.py:108-113 (HF target)
// which the reference computes with three STFTs, two ISTFTs and a dozen elementwise kernels.  For n_fft = 4 the
// STFT is a 3-tap filter bank, so this is pure HBM-bound stencil work: one CTA per (trajectory, channel) row keeps
// the row, its spectrum and the two band-limited reconstructions in shared memory; every output is written once,
// coalesced along time.  Algorithmic bytes per row: 4 L in, 4 (3 * 2 K T + 2 L) out (K = n_fft/2 + 1, T = L/hop + 1).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tvq {

struct FrontendParams {
    const float* x;        // [b, c, l]
    int64_t rows;          // b * c
    int c, l, n_fft;
    float* xf;             // [b, 2c, K, T] or null
    float* enc_in_l;       // [b, 2c, K, T] or null
    float* enc_in_h;       // [b, 2c, K, T] or null
    float* x_l;            // [b, c, l] or null
    float* x_h;            // [b, c, l] or null
};

__global__ void __launch_bounds__(128) frontend_kernel(const FrontendParams p) {
    extern __shared__ float fsm[];
    const int N = p.n_fft, half = N >> 1, hop = N >> 2, K = half + 1;
    const int L = p.l, T = 1 + L / hop, Ly = hop * (T - 1);
    float* xs = fsm;                        // [L + N]   reflect-padded row
    float* xr = xs + ((L + N + 3) & ~3);    // [K][T]    Re X
    float* xi = xr + K * T;                 // [K][T]    Im X
    float* yl = xi + K * T;                 // [Ly]      LF reconstruction
    float* yh = yl + Ly;                    // [Ly]      HF reconstruction
    float* win = yh + Ly;                   // [N]       periodic Hann window
    float* twc = win + N;                   // [K][N]    cos(2 pi k n / N)
    float* tws = twc + K * N;               // [K][N]    sin(2 pi k n / N)
    const int tid = threadIdx.x;
    const float scale = rsqrtf((float)N);
    for (int i = tid; i < N; i += blockDim.x) win[i] = 0.5f - 0.5f * cospif(2.0f * (float)i / (float)N);
    for (int i = tid; i < K * N; i += blockDim.x) {
        const int k = i / N, n = i % N;
        const float a = 2.0f * (float)((k * n) % N) / (float)N;      // exact at the multiples of 1/2 that n_fft = 4 uses
        twc[i] = cospif(a);
        tws[i] = sinpif(a);
    }
    for (int64_t row = blockIdx.x; row < p.rows; row += gridDim.x) {
        __syncthreads();
        const float* xrow = p.x + row * L;
        for (int j = tid; j < L + N; j += blockDim.x) {
            int s = j - half;
            s = s < 0 ? -s : (s >= L ? 2 * (L - 1) - s : s);
            xs[j] = __ldg(xrow + s);
        }
        __syncthreads();
        // ---- STFT: X[k, t] = N^-1/2 sum_n w[n] xp[t hop + n] e^{-2 pi i k n / N}   (t across threads: no index division)
        if (N == 4) {
            // the shipped configuration (configs/config.yaml: n_fft 4): window (0, 1/2, 1, 1/2), twiddles in {0, +-1, +-i}:
            // a 3-tap filter bank, written out
            for (int t = tid; t < T; t += blockDim.x) {
                const float v1 = 0.5f * xs[t + 1], v2 = xs[t + 2], v3 = 0.5f * xs[t + 3];
                xr[t] = 0.5f * ((v1 + v2) + v3);            xi[t] = 0.f;
                xr[T + t] = -0.5f * v2;                     xi[T + t] = 0.5f * (v3 - v1);
                xr[2 * T + t] = 0.5f * ((v2 - v1) - v3);    xi[2 * T + t] = 0.f;
            }
        } else {
            for (int t = tid; t < T; t += blockDim.x) {
                for (int k = 0; k < K; ++k) {
                    float re = 0.f, im = 0.f;
                    for (int n = 0; n < N; ++n) {
                        const float v = win[n] * xs[t * hop + n];
                        re = fmaf(v, twc[k * N + n], re);
                        im = fmaf(-v, tws[k * N + n], im);
                    }
                    xr[k * T + t] = re * scale;
                    xi[k * T + t] = im * scale;
                }
            }
        }
        __syncthreads();
        // ---- spectrogram-shaped outputs, channel = c * 2 + (0 real | 1 imag): [row][z][k][t], coalesced along t
        const int64_t obase = row * 2 * K * T;
        for (int zk = 0; zk < 2 * K; ++zk) {
            const int z = zk >= K, k = zk - z * K;
            const float* src = z ? xi : xr;
            const int kh = (k < 1 ? 1 : k) * T;
            for (int t = tid; t < T; t += blockDim.x) {
                const int64_t o = obase + (int64_t)zk * T + t;
                if (p.xf) p.xf[o] = src[k * T + t];
                if (p.enc_in_l) p.enc_in_l[o] = src[t];                  // bin 0 in every band
                if (p.enc_in_h) p.enc_in_h[o] = src[kh + t];             // bin 1 pasted into bin 0
            }
        }
        // ---- ISTFT of the two band-limited spectra (overlap-add of windowed inverse frames / window envelope)
        if (p.x_l || p.x_h) {
            if (N == 4) {
                // frames t = j + 1, j, j - 1 reach sample j with window taps n = 1, 2, 3 (tap 0 is zero)
                for (int j = tid; j < Ly; j += blockDim.x) {
                    float al = 0.f, ah = 0.f, env = 0.f;
                    if (j + 1 <= T - 1) {                          // n = 1, w = 1/2
                        const int t = j + 1;
                        al += 0.5f * xr[t];
                        ah += 0.5f * (-xr[2 * T + t] - 2.f * xi[T + t]);
                        env += 0.25f;
                    }
                    {                                              // n = 2, w = 1
                        const int t = j;
                        al += xr[t];
                        ah += xr[2 * T + t] - 2.f * xr[T + t];
                        env += 1.f;
                    }
                    if (j >= 1) {                                  // n = 3, w = 1/2
                        const int t = j - 1;
                        al += 0.5f * xr[t];
                        ah += 0.5f * (-xr[2 * T + t] + 2.f * xi[T + t]);
                        env += 0.25f;
                    }
                    yl[j] = al * 0.5f / env;
                    yh[j] = ah * 0.5f / env;
                }
            } else
            for (int j = tid; j < Ly; j += blockDim.x) {
                const int pos = j + half;
                int t0 = (pos - N + hop) / hop;                    // ceil((pos - N + 1) / hop) for pos - N + 1 > 0
                t0 = t0 < 0 ? 0 : t0;
                int t1 = pos / hop;
                t1 = t1 > T - 1 ? T - 1 : t1;
                float al = 0.f, ah = 0.f, env = 0.f;
                for (int t = t0; t <= t1; ++t) {
                    const int n = pos - t * hop;
                    if (n < 0 || n >= N) continue;
                    const float w = win[n];
                    float fh = (n & 1) ? -xr[half * T + t] : xr[half * T + t];
                    for (int k = 1; k < half; ++k)
                        fh += 2.f * (xr[k * T + t] * twc[k * N + n] - xi[k * T + t] * tws[k * N + n]);
                    al = fmaf(w, xr[t], al);
                    ah = fmaf(w, fh, ah);
                    env = fmaf(w, w, env);
                }
                yl[j] = al * scale / env;
                yh[j] = ah * scale / env;
            }
            __syncthreads();
            // ---- F.interpolate(..., size = L, mode = "linear", align_corners = False)
            const float ratio = (float)Ly / (float)L;       // ATen: float scale, source coordinate with ONE rounding (fma)
            for (int j = tid; j < L; j += blockDim.x) {
                float vl, vh;
                if (Ly == L) {
                    vl = yl[j]; vh = yh[j];
                } else {
                    const float s = fmaxf(fmaf(ratio, (float)j + 0.5f, -0.5f), 0.f);
                    int i0 = (int)s;
                    i0 = i0 > Ly - 1 ? Ly - 1 : i0;
                    const int i1 = i0 + 1 > Ly - 1 ? Ly - 1 : i0 + 1;
                    const float lam = s - (float)i0;
                    vl = yl[i0] * (1.f - lam) + yl[i1] * lam;
                    vh = yh[i0] * (1.f - lam) + yh[i1] * lam;
                }
                if (p.x_l) p.x_l[row * L + j] = vl;
                if (p.x_h) p.x_h[row * L + j] = vh;
            }
        }
    }
}

inline size_t frontend_smem_bytes(int l, int n_fft) {
    const int half = n_fft / 2, hop = n_fft / 4, K = half + 1, T = 1 + l / hop, Ly = hop * (T - 1);
    return (size_t)(((l + n_fft + 3) & ~3) + 2 * K * T + 2 * Ly + n_fft + 2 * K * n_fft) * sizeof(float);
}

}  // namespace tvq

// ------------------------------------------------------------------------------------------------------------
// Decoder side (models/vq_vae.py:259-262): y = F.interpolate(timefreq_to_time(pad_func(u)), L) for a spectrogram-
// shaped decoder output u [b, 2c, K, T], where pad_func zeroes every band but one side (zero_pad_high_freq keeps bin
// 0, zero_pad_low_freq keeps bins 1..).  Forward and backward (the op is linear, the backward is its adjoint: an
// STFT-like analysis of g / envelope) as one kernel each; the zeroed bands never leave the kernel, their gradient is
// written as zeros.  band: 0 = all bins (plain timefreq_to_time), 1 = LF (bin 0), 2 = HF (bins >= 1).
// T (frames of u) and L (output length) are independent: the ISTFT yields hop (T - 1) samples, the linear interpolation
// (align_corners = False, no anti-aliasing, ATen's one-rounding source coordinate) maps them onto L — up- or down-sampling
// (the shipped decoders emit T = 384 / 400 frames for L = 200).
namespace tvq {

struct BandIstftParams {
    const float* u;        // forward: [b, 2c, K, T] in;  backward: unused
    const float* g_y;      // backward: [b, c, l] in
    float* y;              // forward: [b, c, l] out
    float* g_u;            // backward: [b, 2c, K, T] out
    int64_t rows;          // b * c
    int l, n_fft, band;
    int t;                 // frames of u (the decoder's own width: 384 / 400 at configs/config.yaml, NOT l / hop + 1)
};

// window envelope sum_t w^2[j + N/2 - t hop] over the frames that exist (torch.istft's normalisation)
__device__ __forceinline__ float istft_envelope(const float* win, int j, int N, int hop, int T) {
    const int pos = j + (N >> 1);
    int t0 = (pos - N + hop) / hop;
    t0 = t0 < 0 ? 0 : t0;
    int t1 = pos / hop;
    t1 = t1 > T - 1 ? T - 1 : t1;
    float env = 0.f;
    for (int t = t0; t <= t1; ++t) {
        const int n = pos - t * hop;
        if (n >= 0 && n < N) env = fmaf(win[n], win[n], env);
    }
    return env;
}

template <bool BACKWARD>
__global__ void __launch_bounds__(128) band_istft_kernel(const BandIstftParams p) {
    extern __shared__ float fsm[];
    const int N = p.n_fft, half = N >> 1, hop = N >> 2, K = half + 1;
    const int L = p.l, T = p.t, Ly = hop * (T - 1);   // torch.istft(center=True) returns hop * (T - 1) samples
    float* xr = fsm;                        // [K][T]
    float* xi = xr + K * T;                 // [K][T]
    float* ys = xi + K * T;                 // [Ly]   signal before the interpolation (forward) / its gradient / envelope (backward)
    float* win = ys + ((Ly + 3) & ~3);      // [N]
    float* twc = win + N;                   // [K][N]
    float* tws = twc + K * N;               // [K][N]
    const int tid = threadIdx.x;
    const float scale = rsqrtf((float)N);
    const int k_lo = p.band == 2 ? 1 : 0, k_hi = p.band == 1 ? 0 : half;     // bins that survive pad_func
    for (int i = tid; i < N; i += blockDim.x) win[i] = 0.5f - 0.5f * cospif(2.0f * (float)i / (float)N);
    for (int i = tid; i < K * N; i += blockDim.x) {
        const int k = i / N, n = i % N;
        const float a = 2.0f * (float)((k * n) % N) / (float)N;
        twc[i] = cospif(a);
        tws[i] = sinpif(a);
    }
    const float ratio = (float)Ly / (float)L;
    for (int64_t row = blockIdx.x; row < p.rows; row += gridDim.x) {
        __syncthreads();
        const int64_t ubase = row * 2 * K * T;
        if (!BACKWARD) {
            for (int zk = 0; zk < 2 * K; ++zk) {
                const int z = zk >= K, k = zk - z * K;
                float* dst = (z ? xi : xr) + k * T;
                const bool keep = k >= k_lo && k <= k_hi;
                for (int t = tid; t < T; t += blockDim.x) dst[t] = keep ? __ldg(p.u + ubase + (int64_t)zk * T + t) : 0.f;
            }
            __syncthreads();
            for (int j = tid; j < Ly; j += blockDim.x) {
                const int pos = j + half;
                int t0 = (pos - N + hop) / hop;
                t0 = t0 < 0 ? 0 : t0;
                int t1 = pos / hop;
                t1 = t1 > T - 1 ? T - 1 : t1;
                float acc = 0.f, env = 0.f;
                for (int t = t0; t <= t1; ++t) {
                    const int n = pos - t * hop;
                    if (n < 0 || n >= N) continue;
                    float f = xr[t] + ((n & 1) ? -xr[half * T + t] : xr[half * T + t]);       // DC and Nyquist: real parts only
                    for (int k = 1; k < half; ++k)
                        f += 2.f * (xr[k * T + t] * twc[k * N + n] - xi[k * T + t] * tws[k * N + n]);
                    acc = fmaf(win[n], f, acc);
                    env = fmaf(win[n], win[n], env);
                }
                ys[j] = acc * scale / env;
            }
            __syncthreads();
            for (int j = tid; j < L; j += blockDim.x) {
                float v;
                if (Ly == L) {
                    v = ys[j];
                } else {
                    const float s = fmaxf(fmaf(ratio, (float)j + 0.5f, -0.5f), 0.f);
                    int i0 = (int)s;
                    i0 = i0 > Ly - 1 ? Ly - 1 : i0;
                    const int i1 = i0 + 1 > Ly - 1 ? Ly - 1 : i0 + 1;
                    const float lam = s - (float)i0;
                    v = ys[i0] * (1.f - lam) + ys[i1] * lam;
                }
                p.y[row * L + j] = v;
            }
        } else {
            // adjoint of the interpolation, then g~[j] = g[j] * scale / envelope[j]
            for (int j = tid; j < Ly; j += blockDim.x) ys[j] = 0.f;
            __syncthreads();
            for (int j = tid; j < L; j += blockDim.x) {
                const float g = __ldg(p.g_y + row * L + j);
                if (Ly == L) {
                    ys[j] = g;
                } else {
                    const float s = fmaxf(fmaf(ratio, (float)j + 0.5f, -0.5f), 0.f);
                    int i0 = (int)s;
                    i0 = i0 > Ly - 1 ? Ly - 1 : i0;
                    const int i1 = i0 + 1 > Ly - 1 ? Ly - 1 : i0 + 1;
                    const float lam = s - (float)i0;
                    atomicAdd(ys + i0, g * (1.f - lam));
                    atomicAdd(ys + i1, g * lam);
                }
            }
            __syncthreads();
            for (int j = tid; j < Ly; j += blockDim.x) ys[j] = ys[j] * scale / istft_envelope(win, j, N, hop, T);
            __syncthreads();
            // g_X[k, t] = c_k sum_n g~[t hop + n - N/2] w[n] (cos, -sin)(2 pi k n / N);  c = 1 for DC / Nyquist (cos only), 2 otherwise
            for (int zk = 0; zk < 2 * K; ++zk) {
                const int z = zk >= K, k = zk - z * K;
                const bool keep = k >= k_lo && k <= k_hi && !(z && (k == 0 || k == half));
                const float ck = (k == 0 || k == half) ? 1.f : 2.f;
                for (int t = tid; t < T; t += blockDim.x) {
                    float acc = 0.f;
                    if (keep) {
                        for (int n = 0; n < N; ++n) {
                            const int j = t * hop + n - half;
                            if (j < 0 || j >= Ly) continue;
                            const float tw = z ? -tws[k * N + n] : twc[k * N + n];
                            acc = fmaf(ys[j] * win[n], tw, acc);
                        }
                        acc *= ck;
                    }
                    p.g_u[ubase + (int64_t)zk * T + t] = acc;
                }
            }
        }
    }
}

inline size_t band_istft_smem_bytes(int t, int n_fft) {
    const int half = n_fft / 2, hop = n_fft / 4, K = half + 1, T = t, Ly = hop * (T - 1);
    return (size_t)(2 * K * T + ((Ly + 3) & ~3) + n_fft + 2 * K * n_fft) * sizeof(float);
}

}  // namespace tvq
