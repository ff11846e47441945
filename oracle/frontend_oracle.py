"""CPU restatement (numpy, float64 internally) of the reference's STFT LF/HF front end — TEST INFRASTRUCTURE ONLY.

Follows /root/reference/timevqvae/utils/train_utils.py:
  time_to_timefreq   :293-307   torch.stft(n_fft, hop = n_fft // 4, periodic Hann window, center=True with reflect
                                padding, onesided, normalized) -> view_as_real -> 'b (c z) n t'
  zero_pad_high_freq :361-372   keep frequency bin 0 (copy=False: zeros elsewhere; copy=True: bin 0 repeated)
  zero_pad_low_freq  :375-386   keep bins 1.. (copy=False: zero bin 0; copy=True: bin 1 pasted into bin 0)
  timefreq_to_time   :310-321   torch.istft (same window, normalized) -> 'b c l'
and the call site /root/reference/timevqvae/trainers/stage1.py:101-113 (F.interpolate(..., input_length, 'linear')).
Pinned against the unmodified reference by tests/golden/frontend_*.npz (oracle/gen_golden_frontend.py).
"""
import numpy as np


def hann(n):
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)          # torch.hann_window(n) (periodic)


def stft(x, n_fft):
    """x (B, C, L) -> xf (B, 2C, n_fft/2+1, T) with channel = c*2 + (0 real | 1 imag)."""
    b, c, l = x.shape
    hop, half = n_fft // 4, n_fft // 2
    xp = np.pad(x.astype(np.float64), ((0, 0), (0, 0), (half, half)), mode="reflect")
    t = 1 + l // hop
    w = hann(n_fft)
    n = np.arange(n_fft)
    k = np.arange(half + 1)
    basis = np.exp(-2j * np.pi * np.outer(k, n) / n_fft) * w[None, :] / np.sqrt(n_fft)       # (K, N)
    frames = np.stack([xp[:, :, i * hop:i * hop + n_fft] for i in range(t)], axis=2)          # (B, C, T, N)
    spec = np.einsum("bctn,kn->bckt", frames, basis)                                         # (B, C, K, T)
    out = np.stack([spec.real, spec.imag], axis=2)                                           # (B, C, 2, K, T)
    return out.reshape(b, 2 * c, half + 1, t)


def zero_pad_high_freq(xf, copy=False):
    if not copy:
        out = np.zeros_like(xf)
        out[:, :, 0, :] = xf[:, :, 0, :]
        return out
    return np.repeat(xf[:, :, [0], :], xf.shape[2], axis=2)


def zero_pad_low_freq(xf, copy=False):
    if not copy:
        out = xf.copy()
        out[:, :, 0, :] = 0
        return out
    return np.concatenate([xf[:, :, [1], :], xf[:, :, 1:, :]], axis=2)


def istft(xf, n_fft):
    """xf (B, 2C, K, T) -> y (B, C, hop*(T-1))."""
    b, c2, kk, t = xf.shape
    c = c2 // 2
    hop, half = n_fft // 4, n_fft // 2
    z = xf.astype(np.float64).reshape(b, c, 2, kk, t)
    spec = (z[:, :, 0] + 1j * z[:, :, 1]) * np.sqrt(n_fft)                                   # undo `normalized`
    frames = np.fft.irfft(spec, n=n_fft, axis=2)                                             # (B, C, N, T)
    w = hann(n_fft)
    total = n_fft + hop * (t - 1)
    y = np.zeros((b, c, total))
    env = np.zeros(total)
    for i in range(t):
        y[:, :, i * hop:i * hop + n_fft] += frames[:, :, :, i] * w[None, None, :]
        env[i * hop:i * hop + n_fft] += w * w
    y, env = y[:, :, half:total - half], env[half:total - half]
    return y / env[None, None, :]


def interp_linear(y, size):
    """F.interpolate(y, size, mode='linear', align_corners=False) along the last axis.  ATen evaluates the source
    coordinate in float32 with ONE rounding, src = fma(in / out, dst + 0.5, -0.5) (clamped at 0); at coordinates of a
    few hundred that rounding (1e-5) is visible in the result, so it is part of the spec (reproduces
    tests/golden/frontend_interp.npz to 2e-7; a float64 coordinate is 3e-5 off, separate float32 roundings 8e-6)."""
    ly = y.shape[-1]
    if ly == size:
        return y.copy()
    scale = np.float64(np.float32(ly) / np.float32(size))
    src = np.maximum((scale * (np.arange(size) + 0.5) - 0.5).astype(np.float32), np.float32(0.0))   # exact product, one rounding
    i0 = np.minimum(src.astype(np.int64), ly - 1)
    i1 = np.minimum(i0 + 1, ly - 1)
    lam = (src - i0.astype(np.float32)).astype(np.float64)
    return y[..., i0] * (1 - lam) + y[..., i1] * lam


def frontend(x, n_fft):
    """Everything stage 1 derives from x before the encoders (trainers/stage1.py:101-113, models/vq_vae.py:179-180)."""
    l = x.shape[-1]
    xf = stft(x, n_fft)
    u_l, u_h = zero_pad_high_freq(xf), zero_pad_low_freq(xf)
    return {"xf": xf.astype(np.float32), "enc_in_l": zero_pad_high_freq(xf, copy=True).astype(np.float32),
            "enc_in_h": zero_pad_low_freq(xf, copy=True).astype(np.float32),
            "x_l": interp_linear(istft(u_l, n_fft), l).astype(np.float32),
            "x_h": interp_linear(istft(u_h, n_fft), l).astype(np.float32)}


def band_istft(u, n_fft, band, length):
    """models/vq_vae.py:259-262: F.interpolate(timefreq_to_time(pad_func(u)), length, 'linear'); band all | lf | hf."""
    v = u if band == "all" else (zero_pad_high_freq(u) if band == "lf" else zero_pad_low_freq(u))
    return interp_linear(istft(v, n_fft), length).astype(np.float32)
