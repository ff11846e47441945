"""Layout glue around the VQ call sites of the reference.

`quantize`  mirrors timevqvae/utils/train_utils.py:338-358 (b c h w <-> b (h w) c around the VQ),
`decode_tokens` mirrors the token -> decoder-input hand-off of timevqvae/models/maskgit.py:465-470,
`lf_hf_frontend` is the whole STFT front end of a stage-1 step in one kernel:
timevqvae/utils/train_utils.py:293-321, :361-386 as used by trainers/stage1.py:101-113 and
models/vq_vae.py:179-180; `time_to_timefreq` keeps the reference's name and signature for its first piece;
`band_timefreq_to_time` is the decoder side (models/vq_vae.py:259-262): pad_func + timefreq_to_time + interpolate,
differentiable, one kernel per direction; `timefreq_to_time` keeps the reference's name for the plain ISTFT.
"""
from __future__ import annotations

from typing import Union

import torch

from . import functional as TF


class _Transpose(torch.autograd.Function):
    """x (b, r, s) -> (b, s, r) contiguous, through the tiled transpose kernel; the backward is the same kernel."""

    @staticmethod
    def forward(ctx, x):
        b, r, s = x.shape
        out = torch.empty(b, s, r, dtype=torch.float32, device=x.device)
        TF._launch("tvq_transpose", x, x.data_ptr(), b, r, s, out.data_ptr())
        return out

    @staticmethod
    def backward(ctx, g):
        return _Transpose.apply(g.contiguous())


def _swap_last_two(x: torch.Tensor) -> torch.Tensor:
    """(b, r, s) -> (b, s, r) contiguous.  fp32 contiguous CUDA input: the tiled kernel; anything else: torch."""
    if x.is_cuda and x.dtype == torch.float32 and x.is_contiguous():
        return _Transpose.apply(x)
    return x.transpose(1, 2).contiguous()


def quantize(z, vq_model, transpose_channel_length_axes: bool = False, svq_temp: Union[float, None] = None):
    """z: (b c h w) or (b c l) encoder output -> (z_q, indices, vq_loss, perplexity)  (utils/train_utils.py:338-358).

    The two layout changes ('b c h w -> b (h w) c' and back) run as coalesced tiled transposes (tvq_transpose);
    z_q comes back contiguous in the input's layout."""
    input_dim = z.dim() - 2
    fused = getattr(vq_model, "_channels_first_ok", None)
    cb = getattr(vq_model, "_codebook", None)
    if fused is not None and vq_model.training and cb._ddp_active() and cb._px is None and z.is_cuda:
        cb.setup_data_parallel(z.device)        # collective, once: every rank's first data-parallel training call
    if fused is not None and (input_dim == 2 or (input_dim == 1 and transpose_channel_length_axes)) and fused(z, svq_temp):
        # the module takes 'b c (h w)' directly: one transpose in, z_q written channels-first, one-kernel backward
        return vq_model.forward_channels_first(z)
    if input_dim == 2:
        b, c, h, w = z.shape
        zz = _swap_last_two(z.reshape(b, c, h * w))                  # b (h w) c
        z_q, indices, vq_loss, perplexity = vq_model(zz, svq_temp)
        z_q = _swap_last_two(z_q).reshape(b, -1, h, w)               # b c h w
    elif input_dim == 1:
        if transpose_channel_length_axes:
            z = _swap_last_two(z)
        z_q, indices, vq_loss, perplexity = vq_model(z, svq_temp)
        if transpose_channel_length_axes:
            z_q = _swap_last_two(z_q)
    else:
        raise ValueError
    return z_q, indices, vq_loss, perplexity


@torch.no_grad()
def decode_tokens(s: torch.Tensor, vq_model, h: int, w: int, *, strict: bool = True) -> torch.Tensor:
    """Token ids (b, n) -> decoder input (b, c, h, w): gather + project_out + 'b n c -> b c h w'.

    Without a projection the gather kernel writes the decoder layout directly (one pass instead
    of the reference's gather + two rearranges).  Ids outside [0, K) raise IndexError (strict, one 4-byte read-back) or
    come out as NaN (strict=False: no synchronisation), as F.embedding would refuse them — never a neighbouring code.
    """
    embed = vq_model._codebook._embed_data()
    s = s.contiguous()
    if isinstance(vq_model.project_out, torch.nn.Identity):
        zq = TF.vq_gather(s, embed, channels_first=True, strict=strict)             # (b, c, n)
    else:
        zq = vq_model.project_out(TF.vq_gather(s, embed, strict=strict)).transpose(1, 2)
    b, c, n = zq.shape
    return zq.reshape(b, c, h, w)


@torch.no_grad()
def lf_hf_frontend(x: torch.Tensor, n_fft: int, *, want=("xf", "enc_in_l", "enc_in_h", "x_l", "x_h")) -> dict:
    """x (b, c, l) fp32 CUDA -> dict of the tensors stage 1 derives from x before the encoders:

    xf        (b, 2c, n_fft/2+1, l/hop+1)  time_to_timefreq(x, n_fft, c)
    enc_in_l  same shape                   zero_pad_high_freq(xf, copy=True)   — LF encoder input
    enc_in_h  same shape                   zero_pad_low_freq(xf, copy=True)    — HF encoder input
    x_l, x_h  (b, c, l)                    F.interpolate(timefreq_to_time(zero_pad_{high,low}_freq(xf), ...), l, "linear")

    One kernel launch; only the entries named in `want` are produced.  No gradient (x is data; the reference
    computes these outside the autograd graph of the parameters as well)."""
    TF._need(x, "x")
    if x.dim() != 3:
        raise ValueError("x must be (b, c, l)")
    b, c, l = x.shape
    hop = n_fft // 4
    if n_fft % 4 or not 4 <= n_fft <= 64 or l <= n_fft // 2:
        raise NotImplementedError(f"n_fft={n_fft}, l={l}: the front-end kernel needs n_fft % 4 == 0, 4 <= n_fft <= 64, l > n_fft / 2")
    k, t = n_fft // 2 + 1, l // hop + 1
    out = {}
    for name in want:
        shape = (b, c, l) if name in ("x_l", "x_h") else (b, 2 * c, k, t)
        out[name] = torch.empty(shape, dtype=torch.float32, device=x.device)
    ptr = lambda name: out[name].data_ptr() if name in out else None
    TF._launch("tvq_frontend", x, x.data_ptr(), b, c, l, n_fft, ptr("xf"), ptr("enc_in_l"), ptr("enc_in_h"), ptr("x_l"), ptr("x_h"))
    return out


def time_to_timefreq(x: torch.Tensor, n_fft: int, C: int, norm: bool = True) -> torch.Tensor:
    """Reference name and signature (utils/train_utils.py:293): x (B, C, L) -> (B, 2C, n_fft/2+1, T)."""
    if not norm:
        raise NotImplementedError("the reference only ever calls time_to_timefreq with norm=True")
    if x.shape[1] != C:
        raise ValueError(f"x has {x.shape[1]} channels, C={C}")
    return lf_hf_frontend(x.contiguous(), n_fft, want=("xf",))["xf"]


class _BandISTFT(torch.autograd.Function):
    """y = F.interpolate(timefreq_to_time(pad_func(u), n_fft, c), length, 'linear'); linear in u, backward = adjoint."""

    @staticmethod
    def forward(ctx, u, n_fft, band, length):
        TF._need(u, "u")
        b, c2, k, t = u.shape
        c = c2 // 2
        y = torch.empty(b, c, length, dtype=torch.float32, device=u.device)
        TF._launch("tvq_band_istft_frames", u, u.data_ptr(), b, c, t, length, n_fft, band, y.data_ptr())
        ctx.meta = (b, c, k, t, n_fft, band, length)
        return y

    @staticmethod
    def backward(ctx, g_y):
        b, c, k, t, n_fft, band, length = ctx.meta
        g_y = g_y.contiguous()
        g_u = torch.empty(b, 2 * c, k, t, dtype=torch.float32, device=g_y.device)
        TF._launch("tvq_band_istft_frames_backward", g_y, g_y.data_ptr(), b, c, t, length, n_fft, band, g_u.data_ptr())
        return g_u, None, None, None


_BANDS = {"all": 0, None: 0, "lf": 1, "LF": 1, "hf": 2, "HF": 2}


def band_timefreq_to_time(u: torch.Tensor, n_fft: int, C: int, band="all", length=None) -> torch.Tensor:
    """Decoder output u (B, 2C, n_fft/2+1, T) -> time series (B, C, length), differentiable in u.

    band "lf": zero_pad_high_freq(u) first (keep bin 0); "hf": zero_pad_low_freq(u) (keep bins 1..); "all": none.
    Equals F.interpolate(timefreq_to_time(pad_func(u), n_fft, C), length, mode="linear") of the reference
    (models/vq_vae.py:259-262) for ANY frame count T (the shipped decoders emit 384 / 400 frames for length 200);
    length defaults to the ISTFT's own length, hop * (T - 1)."""
    if u.dim() != 4 or u.shape[1] != 2 * C or u.shape[2] != n_fft // 2 + 1:
        raise ValueError(f"u must be (B, {2 * C}, {n_fft // 2 + 1}, T), got {tuple(u.shape)}")
    hop = n_fft // 4
    ly = hop * (u.shape[3] - 1)
    length = ly if length is None else int(length)
    if u.shape[3] < 2 or length < 1:
        raise ValueError("need at least two frames and a positive output length")
    return _BandISTFT.apply(u.contiguous().float(), n_fft, _BANDS[band], length)


def timefreq_to_time(x: torch.Tensor, n_fft: int, C: int, norm: bool = True) -> torch.Tensor:
    """Reference name and signature (utils/train_utils.py:310): (B, 2C, N, T) -> (B, C, hop * (T - 1))."""
    if not norm:
        raise NotImplementedError("the reference only ever calls timefreq_to_time with norm=True")
    return band_timefreq_to_time(x, n_fft, C, "all", None)


@torch.no_grad()
def maskgit_step(logits: torch.Tensor, s: torch.Tensor, mask_token_id: int, mask_len: int, temperature: float, *,
                 generator=None, noise=None, return_details: bool = False):
    """One MaskGIT decoding iteration after the transformer (models/maskgit.py:300-346 / :364-410 incl.
    mask_by_random_topk :238-267) in one kernel: logits (b, n, K), s (b, n) int64 -> new s (b, n).

    The noise is drawn from torch's generator exactly as the reference draws it — first the Exp(1) tensor of
    Categorical.sample() (torch.multinomial's one-draw path: argmax(probs / q)), then the U(0,1) tensor of the Gumbel
    perturbation — so the same generator state gives the reference's tokens; pass noise=(q, u) to supply it."""
    TF._need(logits, "logits"); TF._need(s, "s", torch.int64)
    b, n, k = logits.shape
    if noise is None:
        q = torch.empty(b * n, k, dtype=torch.float32, device=logits.device).exponential_(1, generator=generator).view(b, n, k)
        u = torch.zeros(b, n, dtype=torch.float32, device=logits.device).uniform_(0, 1, generator=generator)
    else:
        q, u = noise
        TF._need(q, "q"); TF._need(u, "u")
    s_new = torch.empty_like(s)
    sampled = torch.empty_like(s) if return_details else None
    masking = torch.empty(b, n, dtype=torch.uint8, device=s.device) if return_details else None
    TF._launch("tvq_maskgit_step", logits, logits.data_ptr(), s.data_ptr(), q.data_ptr(), u.data_ptr(), b, n, k, int(mask_token_id),
                                         int(mask_len), float(temperature), s_new.data_ptr(),
                                         sampled.data_ptr() if return_details else None,
                                         masking.data_ptr() if return_details else None)
    return (s_new, sampled, masking.bool()) if return_details else s_new
