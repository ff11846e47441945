"""Stage-1 harness: the caller of the VQ hot path, so that BASELINE's second metric (stage-1 trajectories / second) can be
measured with the new kernels in the loop (VERDICT row N-1).

NOT part of the hot path, and deliberately plain torch where the reference is plain torch: the conv encoder / decoder
stacks restate /root/reference/timevqvae/models/vq_vae.py:13-262 with IDENTICAL module nesting, so the state-dict keys
(`encoder_l.encoder.0.block.0.weight`, `decoder_h.decoder.3.convs.4.bias`, `decoder_l.linear.weight`, ...) and the RNG
consumption order at construction are the reference's: `torch.manual_seed(s); np.random.seed(s); Stage1(...)` yields the
reference's initial weights bit for bit (tests/golden/stage1_*.npz), and its Lightning checkpoints load unchanged.

What is NOT torch here is everything on, or adjacent to, the hot path (include/tvq.h):
  * the STFT LF/HF front end of the step — ONE kernel (tvq_frontend) instead of the reference's three STFTs, two ISTFTs and
    ~10 elementwise kernels (trainers/stage1.py:101-113 and the time_to_timefreq + pad_func at the head of each encoder,
    models/vq_vae.py:179-180);
  * quantize() on both branches — the fused channels-first VQ train step (utils/train_utils.py:338-358, models/vq.py);
  * the decoder tail pad_func -> ISTFT -> interpolate — one kernel per direction (tvq_band_istft, models/vq_vae.py:259-261).

`Stage1.forward` mirrors trainers/stage1.py:89-168 (returns recons_loss, vq_losses, perplexities); `Stage1Trainer` is the
training_step + configure_optimizers of :170-236 (AdamW, linear-warmup cosine schedule) as one CUDA-graph replay per step,
data-parallel with a flat-bucket NCCL all-reduce of the gradients and the EMA statistics exchanged inside the VQ kernels.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn

from . import functional as TF
from . import glue
from .vq import VectorQuantize

__all__ = ["Stage1", "Stage1Trainer", "SnakeActivation", "VQVAEEncoder", "VQVAEDecoder", "compute_downsample_rate",
           "default_config", "warmup_cosine_factor"]


def default_config() -> dict:
    """The keys of /root/reference/configs/config.yaml that stage 1 reads."""
    return {"exp_params": {"lr": 1e-3, "linear_warmup_rate": 0.1},
            "trainer_params": {"max_steps": {"stage1": 50000}},
            "encoder": {"init_dim": 4, "hid_dim": 128, "n_resnet_blocks": 2, "downsampled_width": {"lf": 8, "hf": 32}},
            "decoder": {"n_resnet_blocks": 2},
            "VQ-VAE": {"n_fft": 4, "codebook_sizes": {"lf": 32, "hf": 32}}}


def compute_downsample_rate(input_length: int, n_fft: int, downsampled_width: int) -> int:
    """utils/train_utils.py:413-418."""
    if input_length < downsampled_width:
        return 1
    return round(input_length / (np.log2(n_fft) - 1) / downsampled_width)


class _SnakeFn(torch.autograd.Function):
    """y = x + sin^2(a x) / a on the fused kernels (tvq_snake_forward / _backward): x (b, c, h, w) contiguous in NCHW or
    channels_last order, a (1, c, 1, 1)."""

    @staticmethod
    def forward(ctx, x, a):
        cl = not x.is_contiguous()
        b, c, h, w = x.shape
        y = torch.empty_like(x)
        TF._launch("tvq_snake_forward", x, x.data_ptr(), a.data_ptr(), b, c, h * w, 1 if cl else 0, y.data_ptr())
        ctx.save_for_backward(x, a)
        ctx.cl = cl
        return y

    @staticmethod
    def backward(ctx, g):
        x, a = ctx.saved_tensors
        b, c, h, w = x.shape
        g = g.contiguous(memory_format=torch.channels_last) if ctx.cl else g.contiguous()
        gx = torch.empty_like(x)
        ga = torch.zeros_like(a)
        TF._launch("tvq_snake_backward", x, g.data_ptr(), x.data_ptr(), a.data_ptr(), b, c, h * w, 1 if ctx.cl else 0,
                   gx.data_ptr(), ga.data_ptr())
        return gx, ga


class SnakeActivation(nn.Module):
    """x + sin^2(a x) / a with one learnable `a` per channel (utils/train_utils.py:421-448; 2-D inputs only here).
    `a` is drawn with numpy's global generator, as the reference does — same seed, same parameters.  On CUDA fp32 tensors
    the forward and backward are one kernel each (the reference fuses the expression with TorchScript); anything else
    takes the torch expression."""

    fused = True

    def __init__(self, num_features: int, a_base: float = 0.2, a_max: float = 0.5):
        super().__init__()
        a = np.random.uniform(a_base, a_max, size=(1, num_features, 1, 1))
        self.a = nn.Parameter(torch.tensor(a, dtype=torch.float32))

    def forward(self, x):
        if (SnakeActivation.fused and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and self.a.is_contiguous()
                and (x.is_contiguous() or x.is_contiguous(memory_format=torch.channels_last))):
            return _SnakeFn.apply(x, self.a)
        return x + (1 / self.a) * torch.sin(self.a * x) ** 2


_KS, _PAD = (3, 4), (1, 1)        # strided (de)convolutions of the frequency-dependent variant (frequency_indepence=False)


class ResBlock(nn.Module):
    """models/vq_vae.py:13-62: Snake, 3x3 conv, BN, Snake, 3x3 conv, dropout + (1x1 conv) skip."""

    def __init__(self, cin: int, cout: int, dropout: float = 0.0):
        super().__init__()
        self.convs = nn.Sequential(SnakeActivation(cin), nn.Conv2d(cin, cout, 3, 1, 1), nn.BatchNorm2d(cout),
                                   SnakeActivation(cout), nn.Conv2d(cout, cout, 3, 1, 1), nn.Dropout(dropout))
        self.proj = nn.Identity() if cin == cout else nn.Conv2d(cin, cout, 1)

    def forward(self, x):
        return self.proj(x) + self.convs(x)


class VQVAEEncBlock(nn.Module):
    """models/vq_vae.py:65-92: halves the time axis."""

    def __init__(self, cin: int, cout: int, dropout: float = 0.0):
        super().__init__()
        self.block = nn.Sequential(nn.Conv2d(cin, cout, _KS, (1, 2), _PAD, padding_mode="replicate"), nn.BatchNorm2d(cout),
                                   SnakeActivation(cout), nn.Dropout(dropout))

    def forward(self, x):
        return self.block(x)


class VQVAEDecBlock(nn.Module):
    """models/vq_vae.py:95-121: doubles the time axis."""

    def __init__(self, cin: int, cout: int, dropout: float = 0.0):
        super().__init__()
        self.block = nn.Sequential(nn.ConvTranspose2d(cin, cout, _KS, (1, 2), _PAD), nn.BatchNorm2d(cout),
                                   SnakeActivation(cout), nn.Dropout(dropout))

    def forward(self, x):
        return self.block(x)


def _n_halvings(downsample_rate: int) -> int:
    return int(round(np.log2(downsample_rate)))


class VQVAEEncoder(nn.Module):
    """Conv stack of models/vq_vae.py:124-188.  forward takes the STFT-domain, band-limited input (b, 2c, n_fft/2+1, T) that
    the reference derives inside its forward (:179-180) — here it comes from the one front-end kernel of the step; `encode`
    keeps the reference's (b, c, l) entry point."""

    def __init__(self, init_dim: int, hid_dim: int, num_channels: int, downsample_rate: int, n_resnet_blocks: int,
                 band: str, n_fft: int, dropout: float = 0.3):
        super().__init__()
        self.band, self.n_fft = band, n_fft
        d = init_dim
        layers = [VQVAEEncBlock(num_channels, d)]
        d *= 2
        for _ in range(_n_halvings(downsample_rate) - 1):
            layers.append(VQVAEEncBlock(d // 2, d))
            layers.extend(ResBlock(d, d, dropout) for _ in range(n_resnet_blocks))
            d *= 2
        layers.append(ResBlock(d // 2, hid_dim, dropout))
        self.encoder = nn.Sequential(*layers)
        self.is_num_tokens_updated = False
        self.register_buffer("num_tokens", torch.tensor(0))
        self.register_buffer("H_prime", torch.tensor(0))
        self.register_buffer("W_prime", torch.tensor(0))

    def forward(self, u: torch.Tensor) -> torch.Tensor:
        out = self.encoder(u)
        if not self.is_num_tokens_updated:
            self.H_prime = torch.tensor(out.shape[2])
            self.W_prime = torch.tensor(out.shape[3])
            self.num_tokens = self.H_prime * self.W_prime
            self.is_num_tokens_updated = True
        return out

    def encode(self, x: torch.Tensor) -> torch.Tensor:
        """x (b, c, l) -> z, the reference's VQVAEEncoder.forward."""
        name = "enc_in_l" if self.band == "lf" else "enc_in_h"
        return self(glue.lf_hf_frontend(x.contiguous(), self.n_fft, want=(name,))[name])


class VQVAEDecoder(nn.Module):
    """models/vq_vae.py:191-264; the tail pad_func -> ISTFT -> Upsample runs as one kernel (tvq_band_istft)."""

    def __init__(self, init_dim: int, hid_dim: int, num_channels: int, downsample_rate: int, n_resnet_blocks: int,
                 input_length: int, band: str, n_fft: int, x_channels: int, dropout: float = 0.3):
        super().__init__()
        self.band, self.n_fft, self.x_channels, self.input_length = band, n_fft, x_channels, input_length
        halvings = _n_halvings(downsample_rate)
        d = int(init_dim * 2 ** (halvings - 1)) if halvings != 0 else int(init_dim)
        layers = [ResBlock(hid_dim, d, dropout)]
        for _ in range(halvings - 1):
            layers.extend(ResBlock(d, d, dropout) for _ in range(n_resnet_blocks))
            d //= 2
            layers.append(VQVAEDecBlock(2 * d, d))
        layers.append(nn.ConvTranspose2d(d, num_channels, _KS, (1, 2), _PAD))
        layers.append(nn.ConvTranspose2d(num_channels, num_channels, _KS, (1, 2), _PAD))
        self.decoder = nn.Sequential(*layers)
        self.interp = nn.Upsample(input_length, mode="linear")
        self.linear = nn.Linear(input_length, input_length)

    def forward(self, zq: torch.Tensor) -> torch.Tensor:
        u = self.decoder(zq)                   # (b, 2c, n_fft/2+1, T'): T' = 384 (LF) / 400 (HF) frames at configs/config.yaml
        out = glue.band_timefreq_to_time(u, self.n_fft, self.x_channels, self.band, self.input_length)
        return out + self.linear(out)


class Stage1(nn.Module):
    """trainers/stage1.py:15-87 (construction) and :89-168 (forward) without Lightning."""

    def __init__(self, input_length: int, in_channels: int, config: dict, **kwargs):
        super().__init__()
        self.input_length, self.in_channels, self.config = input_length, in_channels, config
        self.n_fft = config["VQ-VAE"]["n_fft"]
        enc, init_dim, hid_dim = config["encoder"], config["encoder"]["init_dim"], config["encoder"]["hid_dim"]
        rate_l = compute_downsample_rate(input_length, self.n_fft, enc["downsampled_width"]["lf"])
        rate_h = compute_downsample_rate(input_length, self.n_fft, enc["downsampled_width"]["hf"])
        c2 = 2 * in_channels
        # construction order = the reference's (it fixes which random numbers each layer receives)
        self.encoder_l = VQVAEEncoder(init_dim, hid_dim, c2, rate_l, enc["n_resnet_blocks"], "lf", self.n_fft)
        self.encoder_h = VQVAEEncoder(init_dim, hid_dim, c2, rate_h, enc["n_resnet_blocks"], "hf", self.n_fft)
        self.vq_model_l = VectorQuantize(hid_dim, config["VQ-VAE"]["codebook_sizes"]["lf"], **config["VQ-VAE"])
        self.vq_model_h = VectorQuantize(hid_dim, config["VQ-VAE"]["codebook_sizes"]["hf"], **config["VQ-VAE"])
        nres = config["decoder"]["n_resnet_blocks"]
        self.decoder_l = VQVAEDecoder(init_dim, hid_dim, c2, rate_l, nres, input_length, "lf", self.n_fft, in_channels)
        self.decoder_h = VQVAEDecoder(init_dim, hid_dim, c2, rate_h, nres, input_length, "hf", self.n_fft, in_channels)

    # The LF and HF branches (encoder -> quantize -> decoder -> loss) are independent until the losses are added: on CUDA
    # they run on two streams (autograd replays each branch's backward on the stream of its forward), so the many small
    # kernels of the deep, narrow layers overlap.  Set to False for one stream.
    two_streams = True

    def _branch(self, enc, vq, dec, u, target, loss_fn):
        z = enc(u)
        z_q, s, vq_loss, ppl = glue.quantize(z, vq)
        xhat = dec(z_q)
        return xhat, (loss_fn(target, xhat) if target is not None else None), vq_loss, ppl

    def forward(self, batch, batch_idx: int = 0, return_x_rec: bool = False):
        x, _ = batch if isinstance(batch, (tuple, list)) else (batch, None)
        front = glue.lf_hf_frontend(x.contiguous(), self.n_fft, want=("enc_in_l", "enc_in_h", "x_l", "x_h"))
        tl, th = (None, None) if return_x_rec else (front["x_l"], front["x_h"])
        if self.two_streams and x.is_cuda:
            cur = torch.cuda.current_stream(x.device)
            if getattr(self, "_side_stream", None) is None:
                self._side_stream = torch.cuda.Stream(x.device)
            side = self._side_stream
            side.wait_stream(cur)
            for t in (front["enc_in_h"], front["x_h"]):
                t.record_stream(side)
            with torch.cuda.stream(side):
                xhat_h, loss_h, vq_loss_h, ppl_h = self._branch(self.encoder_h, self.vq_model_h, self.decoder_h, front["enc_in_h"], th, F.l1_loss)
            xhat_l, loss_l, vq_loss_l, ppl_l = self._branch(self.encoder_l, self.vq_model_l, self.decoder_l, front["enc_in_l"], tl, F.mse_loss)
            cur.wait_stream(side)
            for t in (xhat_h, loss_h, vq_loss_h["loss"], ppl_h):
                if torch.is_tensor(t):
                    t.record_stream(cur)
        else:
            xhat_l, loss_l, vq_loss_l, ppl_l = self._branch(self.encoder_l, self.vq_model_l, self.decoder_l, front["enc_in_l"], tl, F.mse_loss)
            xhat_h, loss_h, vq_loss_h, ppl_h = self._branch(self.encoder_h, self.vq_model_h, self.decoder_h, front["enc_in_h"], th, F.l1_loss)
        if return_x_rec:
            return xhat_l + xhat_h
        recons_loss = {"LF.time": loss_l, "HF.time": loss_h}
        return recons_loss, {"LF": vq_loss_l, "HF": vq_loss_h}, {"LF": ppl_l, "HF": ppl_h}

    def total_loss(self, batch) -> Dict[str, torch.Tensor]:
        """The loss dictionary of training_step (trainers/stage1.py:170-198)."""
        recons, vql, ppl = self.forward(batch)
        loss = (recons["LF.time"] + recons["HF.time"]) + vql["LF"]["loss"] + vql["HF"]["loss"]
        return {"loss": loss, "recons_loss.time": recons["LF.time"] + recons["HF.time"],
                "recons_loss.LF.time": recons["LF.time"], "recons_loss.HF.time": recons["HF.time"],
                "commit_loss.LF": vql["LF"]["commit_loss"], "commit_loss.HF": vql["HF"]["commit_loss"],
                "perplexity.LF": ppl["LF"], "perplexity.HF": ppl["HF"]}


def warmup_cosine_factor(step: int, max_steps: int, warmup_rate: float = 0.1, base_lr: float = 1e-3, min_lr: float = 1e-6) -> float:
    """lr(step) / base_lr of linear_warmup_cosine_annealingLR (utils/train_utils.py:451-472): LambdaLR warm-up chained
    with CosineAnnealingLR(T_max = max_steps - warmup, eta_min = min_lr) at the milestone."""
    warm = int(max_steps * warmup_rate)
    if step < warm:
        return float(step) / float(max(1, warm))
    t, t_max = step - warm, max(1, max_steps - warm)
    return (min_lr + (base_lr - min_lr) * (1 + math.cos(math.pi * t / t_max)) / 2) / base_lr


class Stage1Trainer:
    """One stage-1 optimisation step — forward, backward, (gradient all-reduce), AdamW — captured ONCE as a CUDA graph and
    replayed per batch (trainers/stage1.py:170-236, scripts/train.py:29-43 without Lightning).

    Data parallel (an initialised NCCL group): the batch is sharded by the caller, the EMA statistics of both codebooks
    are exchanged inside the VQ kernels (sync_codebook=True), and the gradients of all parameters live in ONE flat
    buffer that is all-reduced (NCCL, averaged) inside the graph before the optimizer step — the reference's own
    precedent is `devices=1` (scripts/train.py:38); this is the DDP the BASELINE configs[3] asks for."""

    def __init__(self, model: Stage1, batch_shape, *, lr: Optional[float] = None, use_graph: bool = True, group=None):
        self.model = model.train()
        dev = next(model.parameters()).device
        self.device = dev
        cfg = model.config
        self.base_lr = float(lr if lr is not None else cfg["exp_params"]["lr"])
        self.max_steps = int(cfg["trainer_params"]["max_steps"]["stage1"])
        self.warmup_rate = float(cfg["exp_params"]["linear_warmup_rate"])
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.group = group
        params = [p for p in model.parameters() if p.requires_grad]
        self.flat_grad = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=dev)
        off = 0
        for p in params:                      # every .grad is a view of the flat bucket: one all-reduce, no copies
            # same sizes AND strides as the parameter (a channels_last weight is a dense permutation), as autograd's
            # gradient-layout contract and the fused optimizer want
            p.grad = self.flat_grad[off:off + p.numel()].as_strided(p.size(), p.stride())
            off += p.numel()
        self.lr = torch.tensor(self.base_lr * warmup_cosine_factor(1, self.max_steps, self.warmup_rate, self.base_lr),
                               dtype=torch.float32, device=dev)
        self.opt = torch.optim.AdamW(params, lr=self.lr, capturable=True, fused=True)
        self.x = torch.zeros(batch_shape, dtype=torch.float32, device=dev)       # static input of the graph
        self.out: Dict[str, torch.Tensor] = {}
        self.step_count = 0
        self.graph = None
        self.use_graph = use_graph

    def _step_body(self):
        self.flat_grad.zero_()
        out = self.model.total_loss((self.x, None))
        out["loss"].sum().backward()
        for vq in (self.model.vq_model_l, self.model.vq_model_h):      # deferred EMA exchange (if enabled): rejoin this stream
            vq._codebook.join_pending()
        if self.world > 1:
            dist.all_reduce(self.flat_grad, group=self.group)
            self.flat_grad.mul_(1.0 / self.world)
        self.opt.step()
        return {k: (v.detach() if torch.is_tensor(v) else torch.tensor(float(v), device=self.device)) for k, v in out.items()}

    def warmup_and_capture(self, warmup_steps: int = 3):
        """Eager warm-up steps on a side stream (cuDNN autotune, optimizer state, peer-exchange rendezvous), then capture."""
        s = torch.cuda.Stream(self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(warmup_steps):
                self.out = self._step_body()
                self.step_count += 1
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        if not self.use_graph:
            return
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.out = self._step_body()
        torch.cuda.synchronize(self.device)

    def step(self, x: torch.Tensor, non_blocking: bool = True) -> Dict[str, torch.Tensor]:
        """One optimisation step on batch x (host or device tensor, this rank's shard)."""
        self.step_count += 1
        self.lr.fill_(self.base_lr * warmup_cosine_factor(self.step_count, self.max_steps, self.warmup_rate, self.base_lr))
        self.x.copy_(x, non_blocking=non_blocking)
        if self.graph is not None:
            self.graph.replay()
        else:
            self.out = self._step_body()
        return self.out
