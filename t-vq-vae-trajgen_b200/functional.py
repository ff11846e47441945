"""Tensor-level wrappers of the C ABI (include/tvq.h) and the autograd step built on them.

PyTorch is plumbing here: it owns the device memory and the stream; every arithmetic step of
the hot path runs in libtvq_b200.so.  All functions need contiguous fp32 CUDA tensors and raise
otherwise — there is no CPU or eager fallback.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib

__all__ = ["Workspace", "vq_forward_raw", "vq_train_step_raw", "vq_ema_update", "PeerExchange", "vq_ema_update_dp", "vq_backward", "vq_gather", "vq_neg_dist",
           "vq_reseed", "stats_offset", "stats_len", "VQTrainStep", "VQTrainStepCF", "transpose12", "vq_forward_qcf", "vq_forward_cf"]


def stats_offset(k: int) -> int:
    """Element offset of embed_sum inside the packed statistics buffer (TVQ_STATS_OFFSET)."""
    return (k + 3) & ~3


def stats_len(k: int, d: int) -> int:
    return stats_offset(k) + k * d


def _stream(device=None) -> int:
    """Raw cudaStream_t of torch's current stream on `device` (default: the current device)."""
    return torch.cuda.current_stream(device).cuda_stream


class _on_device_of:
    """Make `t`'s device the current CUDA device for the enclosed C-ABI call.  The ABI works on the CURRENT device
    (include/tvq.h), so a module living on cuda:1 must not launch on cuda:0 just because nobody called set_device."""
    __slots__ = ("idx", "prev")

    def __init__(self, t: torch.Tensor):
        self.idx = t.device.index

    def __enter__(self):
        self.prev = torch.cuda.current_device()
        if self.prev != self.idx:
            torch.cuda.set_device(self.idx)
        return self

    def __exit__(self, *exc):
        if self.prev != self.idx:
            torch.cuda.set_device(self.prev)
        return False


def _launch(name: str, anchor: torch.Tensor, *args) -> None:
    """lib.<name>(*args, stream) on the anchor tensor's device and torch's current stream there; raises on a non-zero
    status.  Every pointer argument must live on that device (checked by the callers through _same_device)."""
    with _on_device_of(anchor):
        rc = getattr(_lib.load(), name)(*args, torch.cuda.current_stream(anchor.device).cuda_stream)
    _lib.check(rc, name)


def _same_device(anchor: torch.Tensor, *others) -> None:
    for t in others:
        if t is not None and t.device != anchor.device:
            raise RuntimeError(f"tensors on different devices: {anchor.device} and {t.device}")


def _need(t: torch.Tensor, name: str, dtype=torch.float32) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name}: the B200 VQ kernels need a CUDA tensor (got "
                           f"{getattr(t, 'device', type(t))}); there is no CPU fallback")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous tensor")


class Workspace:
    """Per-codebook device scratch: launch-ticket header + |e|^2 table, and the packed statistics."""

    def __init__(self, k: int, d: int, device: torch.device):
        lib = _lib.load()
        self.k, self.d, self.device = k, d, device
        self.nbytes = int(lib.tvq_workspace_bytes(0, k, d))
        self.buf = torch.zeros(self.nbytes, dtype=torch.uint8, device=device)   # zeroed once (ticket)
        self.stats = torch.empty(stats_len(k, d), dtype=torch.float32, device=device)

    def matches(self, k: int, d: int, device: torch.device) -> bool:
        return self.k == k and self.d == d and self.device == device


def vq_forward_raw(x: torch.Tensor, codebook: torch.Tensor, ws: Workspace, *, train: bool, write_q: bool = True,
                   idx: Optional[torch.Tensor] = None, flags: int = 0, commitment_weight: float = 1.0
                   ) -> Tuple[torch.Tensor, Optional[torch.Tensor], torch.Tensor]:
    """tvq_forward: (idx[n] int64, q[n,d] or None, scalars[8]).  Statistics land in ws.stats.

    `idx` given -> the codes are an input (TVQ_F_GIVEN_IDX) and only gather/ST/loss/statistics run.
    scalars: [0] commit loss (train), [1] perplexity, [2] commitment_weight * commit loss,
    [4:6] uint32 diagnostics (view as int32): rows re-scored in fp64, rows fully re-scanned.
    """
    _need(x, "x")
    _need(codebook, "codebook")
    _same_device(x, codebook, ws.buf, idx)
    n, d = x.shape
    k = codebook.shape[0]
    if codebook.shape[1] != d:
        raise ValueError(f"codebook dim {codebook.shape[1]} != latent dim {d}")
    f = flags | (_lib.F_TRAIN if train else 0) | (_lib.F_WRITE_Q if write_q else 0)
    if idx is None:
        idx = torch.empty(n, dtype=torch.int64, device=x.device)
    else:
        _need(idx, "idx", torch.int64)
        f |= _lib.F_GIVEN_IDX
    q = torch.empty_like(x) if write_q else None
    scalars = torch.empty(_lib.NUM_SCALARS, dtype=torch.float32, device=x.device)
    if n == 0:                          # nothing to launch: zero statistics, NaN means (as the reference)
        scalars.fill_(float("nan"))
        ws.stats.zero_()
        return idx, q, scalars
    _launch("tvq_forward", x, x.data_ptr(), codebook.data_ptr(), n, k, d, f, float(commitment_weight), idx.data_ptr(),
                                 q.data_ptr() if write_q else None, ws.stats.data_ptr(), scalars.data_ptr(),
                                 ws.buf.data_ptr(), ws.nbytes)
    return idx, q, scalars


def vq_ema_update(stats: torch.Tensor, cluster_size: torch.Tensor, embed_avg: torch.Tensor, embed: torch.Tensor,
                  embed_prev: Optional[torch.Tensor], decay: float, eps: float, ws: Workspace) -> None:
    _need(stats, "stats"); _need(cluster_size, "cluster_size"); _need(embed_avg, "embed_avg"); _need(embed, "embed")
    k, d = embed.shape
    _launch("tvq_ema_update", stats, stats.data_ptr(), cluster_size.data_ptr(), embed_avg.data_ptr(), embed.data_ptr(),
                                    embed_prev.data_ptr() if embed_prev is not None else None, k, d, float(decay),
                                    float(eps), ws.buf.data_ptr(), ws.nbytes)


class PeerExchange:
    """Per-codebook exchange buffers for the fused NVLink all-reduce + EMA kernel (tvq_ema_update_dp).

    Every rank allocates the same buffer with torch's symmetric-memory allocator; the rendezvous maps
    all peers' buffers into this process, and the device array of peer pointers is what the kernel
    takes.  PyTorch is plumbing here (allocation + handle exchange); the exchange itself — remote
    stores, flags, the rank-ordered sum — is in the kernel.  Raises if symmetric memory is not
    available; the caller then keeps the NCCL all-reduce path."""

    def __init__(self, k: int, d: int, device: torch.device, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        lib = _lib.load()
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.k, self.d, self.device = k, d, device
        nbytes = int(lib.tvq_exchange_bytes(k, d, self.world))
        if nbytes == 0 or stats_len(k, d) > (1 << 16):
            raise ValueError("statistics too large for the one-CTA exchange kernel")
        self.buf = symm.empty(nbytes, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, group)
        self.peers = torch.tensor([int(p) for p in self.handle.buffer_ptrs], dtype=torch.int64, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group)                 # every rank's buffer is zeroed and mapped before the first step

    def matches(self, k: int, d: int, device: torch.device) -> bool:
        return self.k == k and self.d == d and self.device == device

    def check(self) -> None:
        """Raise if a kernel ever gave up waiting for a peer (error word of the exchange header, include/tvq.h)."""
        step = int(self.buf[4:8].view(torch.int32).item())
        if step != 0:
            raise RuntimeError(f"tvq_b200: rank {self.rank} timed out waiting for a peer's codebook statistics in "
                               f"data-parallel step {step}; the replicas' codebooks may have diverged")


def vq_ema_update_dp(stats: torch.Tensor, ex: PeerExchange, cluster_size: torch.Tensor, embed_avg: torch.Tensor,
                     embed: torch.Tensor, embed_prev: Optional[torch.Tensor], decay: float, eps: float) -> None:
    """tvq_ema_update_dp: all-reduce of the packed statistics over NVLink peer memory + EMA update, one kernel."""
    _need(stats, "stats"); _need(cluster_size, "cluster_size"); _need(embed_avg, "embed_avg"); _need(embed, "embed")
    k, d = embed.shape
    _launch("tvq_ema_update_dp", stats, stats.data_ptr(), ex.peers.data_ptr(), ex.rank, ex.world, cluster_size.data_ptr(),
                                       embed_avg.data_ptr(), embed.data_ptr(),
                                       embed_prev.data_ptr() if embed_prev is not None else None, k, d, float(decay),
                                       float(eps))


_SM_COUNT = {}


def _hint_sm_share(cb, device) -> None:
    """EuclideanCodebook.sm_share -> tvq_hint_max_ctas for the launch that follows (same thread)."""
    share = getattr(cb, "sm_share", None)
    if share:
        sms = _SM_COUNT.get(device.index)
        if sms is None:
            sms = _SM_COUNT[device.index] = torch.cuda.get_device_properties(device).multi_processor_count
        _lib.load().tvq_hint_max_ctas(max(1, int(round(float(share) * sms))))


def _defer_begin(cb, px) -> int:
    """Deferred data-parallel exchange (EuclideanCodebook.defer_exchange): one-shot hint for the launch that follows."""
    mode = int(getattr(cb, "defer_exchange", 0) or 0) if px is not None else 0
    if mode:
        mode = 2 if mode is True or mode == 2 else 1        # True: everything after the scalars leaves the forward kernel
        _lib.load().tvq_hint_defer_exchange(mode)
    return mode


def _defer_finish(cb, px, anchor: torch.Tensor, mode: int, ws: "Workspace") -> None:
    """Enqueue tvq_ema_finalize_dp on the codebook's side stream, ordered after the train step just launched on the current
    stream; the event it records is what every later reader of embed / embed_avg / cluster_size waits for
    (EuclideanCodebook.join_pending — the module's buffer accessors do it)."""
    dev = anchor.device
    cur = torch.cuda.current_stream(dev)
    if cb._side is None:
        cb._side = torch.cuda.Stream(dev)
    ev = torch.cuda.Event()
    ev.record(cur)
    cb._side.wait_event(ev)
    embed = cb._embed_data()
    k, d = embed.shape
    with torch.cuda.stream(cb._side):
        _launch("tvq_ema_finalize_dp", anchor, 1 if mode == 1 else 0, ws.buf.data_ptr(), ws.nbytes, px.peers.data_ptr(), px.rank,
                px.world, cb._buffers["cluster_size"].data_ptr(),
                cb._buffers["embed_avg"].data_ptr(), embed.data_ptr(), k, d, float(cb.decay), float(cb.eps))
        done = torch.cuda.Event()
        done.record(cb._side)
    cb.__dict__["_pending"] = done


def vq_train_step_raw(x: torch.Tensor, cb, ws: Workspace, commitment_weight: float, embed_prev: Optional[torch.Tensor],
                      px: Optional["PeerExchange"] = None):
    """tvq_train_step on a codebook module's buffers: fused forward + EMA (one kernel for k <= 32, d <= 128).

    Returns (idx, q_st, scalars[8], commit[()], weighted[1]); the module's cluster_size / embed_avg /
    embed are updated in place, `embed_prev` (optional) receives the codebook the outputs came from.
    """
    _need(x, "x")
    embed = cb._embed_data()
    _same_device(x, embed, ws.buf, embed_prev)
    _hint_sm_share(cb, x.device)
    n, d = x.shape
    k = embed.shape[0]
    idx = torch.empty(n, dtype=torch.int64, device=x.device)
    q = torch.empty_like(x)
    scalars = torch.empty(_lib.NUM_SCALARS, dtype=torch.float32, device=x.device)
    commit = torch.empty((), dtype=torch.float32, device=x.device)
    weighted = torch.empty(1, dtype=torch.float32, device=x.device)
    if n == 0:
        scalars.fill_(float("nan")); commit.fill_(float("nan")); weighted.fill_(float("nan"))
    if px is not None:       # data-parallel: the kernel's last CTA sums the statistics of all ranks over NVLink peer memory
        deferred = _defer_begin(cb, px)
        _launch("tvq_train_step_dp", x, x.data_ptr(), embed.data_ptr(), cb.cluster_size.data_ptr(), cb.embed_avg.data_ptr(),
                                           embed_prev.data_ptr() if embed_prev is not None else None, n, k, d,
                                           float(commitment_weight), float(cb.decay), float(cb.eps), idx.data_ptr(),
                                           q.data_ptr(), scalars.data_ptr(), commit.data_ptr(), weighted.data_ptr(),
                                           ws.buf.data_ptr(), ws.nbytes, px.peers.data_ptr(), px.rank, px.world)
        if deferred:
            _defer_finish(cb, px, x, deferred, ws)
        return idx, q, scalars, commit, weighted
    _launch("tvq_train_step", embed, x.data_ptr() if n else None, embed.data_ptr(), cb.cluster_size.data_ptr(),
                                    cb.embed_avg.data_ptr(), embed_prev.data_ptr() if embed_prev is not None else None,
                                    n, k, d, float(commitment_weight), float(cb.decay), float(cb.eps),
                                    idx.data_ptr() if n else None, q.data_ptr() if n else None, scalars.data_ptr(),
                                    commit.data_ptr(), weighted.data_ptr(), ws.buf.data_ptr(), ws.nbytes)
    return idx, q, scalars, commit, weighted


def vq_backward(g_q: Optional[torch.Tensor], g_commit: Optional[torch.Tensor], g_weighted: Optional[torch.Tensor],
                x: torch.Tensor, idx: torch.Tensor, codebook: torch.Tensor, commitment_weight: float) -> torch.Tensor:
    """g_x = g_q + (g_commit + w * g_weighted) * 2/(n d) * (x - q_st); any gradient may be None (= 0)."""
    _need(x, "x"); _need(idx, "idx", torch.int64); _need(codebook, "codebook")
    for t, name in ((g_q, "g_q"), (g_commit, "g_commit"), (g_weighted, "g_weighted")):
        if t is not None:
            _need(t, name)
    _same_device(x, idx, codebook, g_q, g_commit, g_weighted)
    n, d = x.shape
    g_x = torch.empty_like(x)
    ptr = lambda t: t.data_ptr() if t is not None else None
    _launch("tvq_backward", x, ptr(g_q), ptr(g_commit), ptr(g_weighted), x.data_ptr(), idx.data_ptr(),
                                  codebook.data_ptr(), n, codebook.shape[0], d, float(commitment_weight), g_x.data_ptr())
    return g_x


def vq_gather(tokens: torch.Tensor, codebook: torch.Tensor, channels_first: bool = False, *, strict: bool = True
              ) -> torch.Tensor:
    """tokens (b, t) int64 -> (b, t, d), or (b, d, t) when channels_first (models/maskgit.py:465-470).

    An id outside [0, k) — e.g. a mask token (id == k) left by an incomplete MaskGIT pass — is an error, as it is for the
    reference's F.embedding: the kernel writes NaN for that token and counts it; strict=True (default) reads the count
    (one 4-byte device-to-host read) and raises IndexError, strict=False leaves the NaNs to speak for themselves."""
    _need(tokens, "tokens", torch.int64)
    _need(codebook, "codebook")
    _same_device(tokens, codebook)
    if tokens.dim() != 2:
        raise ValueError("tokens must be (b, t)")
    b, t = tokens.shape
    k, d = codebook.shape
    out = torch.empty((b, d, t) if channels_first else (b, t, d), dtype=torch.float32, device=tokens.device)
    bad = torch.zeros(1, dtype=torch.int32, device=tokens.device) if strict else None
    _launch("tvq_gather_checked", tokens, tokens.data_ptr(), codebook.data_ptr(), b, t, k, d, 1 if channels_first else 0,
            out.data_ptr(), bad.data_ptr() if strict else None)
    if strict and int(bad.item()) != 0:
        raise IndexError(f"vq_gather: {int(bad.item())} token id(s) outside [0, {k}) (a leaked mask token?)")
    return out


def vq_neg_dist(x: torch.Tensor, codebook: torch.Tensor) -> torch.Tensor:
    """Dense -(|x|^2 - 2 x.e + |e|^2) [n, k] for the stochastic branch (vq.py:210-214)."""
    _need(x, "x"); _need(codebook, "codebook")
    n, d = x.shape
    k = codebook.shape[0]
    dist = torch.empty((n, k), dtype=torch.float32, device=x.device)
    _launch("tvq_neg_dist", x, x.data_ptr(), codebook.data_ptr(), n, k, d, dist.data_ptr())
    return dist


def vq_reseed(x: torch.Tensor, rows: torch.Tensor, cluster_size: torch.Tensor, threshold: float,
              embed: torch.Tensor) -> None:
    _need(x, "x"); _need(rows, "rows", torch.int64); _need(cluster_size, "cluster_size"); _need(embed, "embed")
    n, d = x.shape
    _launch("tvq_reseed", x, x.data_ptr(), rows.data_ptr(), cluster_size.data_ptr(), float(threshold),
                                embed.data_ptr(), n, embed.shape[0], d)


class VQTrainStep(torch.autograd.Function):
    """One training-mode codebook step: assign + ST + commit loss + EMA, differentiable in x.

    forward(x[n,d], codebook_module, commitment_weight, given_idx|None)
        -> (q_st[n,d], idx[n], scalars[8], commit[()], weighted[1])
           scalars[1] = perplexity; commit = mean((q_st - x)^2); weighted = commitment_weight * commit
    backward: g_x = g_q + (g_commit + w g_weighted) * 2/(n d) * (x - q_st)          

    Without data-parallel statistics the whole step is tvq_train_step (one launch where the shape
    allows).  With `sync_codebook` on an initialised process group the packed statistics are
    all-reduced between the forward and the EMA kernel (vq.py:229/234).  The pre-update codebook is
    kept for the backward.
    """

    @staticmethod
    def forward(ctx, x, cb, commitment_weight, given_idx):
        ctx.set_materialize_grads(False)
        ws = cb._workspace(x.device)
        prev = torch.empty_like(cb._embed_data()) if ctx.needs_input_grad[0] else None
        px = None
        fused_dp = False
        if given_idx is None and cb._ddp_active() and cb.codebook_size <= 32 and cb.dim <= 128 and x.shape[0] >= 1:
            px = cb._peer_exchange(x.device)
            fused_dp = px is not None
        if given_idx is None and (fused_dp or not cb._ddp_active()):
            idx, q, scalars, commit, weighted = vq_train_step_raw(x, cb, ws, commitment_weight, prev, px)
        else:
            idx, q, scalars = vq_forward_raw(x, cb._embed_data(), ws, train=True, write_q=True, idx=given_idx,
                                             commitment_weight=commitment_weight)
            cb._sync_and_update(ws, prev)
            commit, weighted = scalars[0].clone(), scalars[2:3].clone()
        ctx.save_for_backward(x, idx, prev)
        ctx.commitment_weight = float(commitment_weight)
        ctx.mark_non_differentiable(idx, scalars)
        return q, idx, scalars, commit, weighted

    @staticmethod
    def backward(ctx, g_q, g_idx, g_scalars, g_commit, g_weighted):
        x, idx, prev = ctx.saved_tensors
        if g_q is not None and not g_q.is_contiguous():
            g_q = g_q.contiguous()
        if g_commit is None and g_weighted is None:
            return g_q, None, None, None
        g_commit = g_commit.contiguous() if g_commit is not None else None
        g_weighted = g_weighted.contiguous() if g_weighted is not None else None
        return vq_backward(g_q, g_commit, g_weighted, x, idx, prev, ctx.commitment_weight), None, None, None


# ------------------------------------------------------------------------------------------------------------------
# Channels-first call site: quantize() hands over z as 'b c (h w)' and wants z_q back the same way.

def transpose12(x: torch.Tensor) -> torch.Tensor:
    """(b, r, s) fp32 contiguous CUDA -> (b, s, r) contiguous through the tiled transpose kernel (no autograd)."""
    _need(x, "x")
    b, r, s = x.shape
    out = torch.empty(b, s, r, dtype=torch.float32, device=x.device)
    _launch("tvq_transpose", x, x.data_ptr(), b, r, s, out.data_ptr())
    return out


def vq_forward_qcf(x: torch.Tensor, codebook: torch.Tensor, ws: Workspace, hw: int, *, train: bool,
                   commitment_weight: float = 1.0):
    """tvq_forward_qcf: x [n, d] row-major -> (idx [n], q [n / hw, d, hw] channels-first, scalars[8])."""
    _need(x, "x"); _need(codebook, "codebook")
    n, d = x.shape
    k = codebook.shape[0]
    idx = torch.empty(n, dtype=torch.int64, device=x.device)
    q = torch.empty(n // hw, d, hw, dtype=torch.float32, device=x.device)
    scalars = torch.empty(_lib.NUM_SCALARS, dtype=torch.float32, device=x.device)
    f = (_lib.F_TRAIN if train else 0) | _lib.F_WRITE_Q
    _launch("tvq_forward_qcf", x, x.data_ptr(), codebook.data_ptr(), n, k, d, f, float(commitment_weight), idx.data_ptr(),
                                     q.data_ptr(), ws.stats.data_ptr(), scalars.data_ptr(), ws.buf.data_ptr(), ws.nbytes, int(hw))
    return idx, q, scalars


def vq_forward_cf(z: torch.Tensor, codebook: torch.Tensor, ws: Workspace, *, train: bool, write_q: bool = True,
                  commitment_weight: float = 1.0):
    """tvq_forward_cf: z [b, d, hw] channels-first, read in place -> (idx [b * hw], q [b, d, hw] or None, scalars[8])."""
    _need(z, "z"); _need(codebook, "codebook")
    b, d, hw = z.shape
    k = codebook.shape[0]
    idx = torch.empty(b * hw, dtype=torch.int64, device=z.device)
    q = torch.empty(b, d, hw, dtype=torch.float32, device=z.device) if write_q else None
    scalars = torch.empty(_lib.NUM_SCALARS, dtype=torch.float32, device=z.device)
    f = (_lib.F_TRAIN if train else 0) | (_lib.F_WRITE_Q if write_q else 0)
    _launch("tvq_forward_cf", z, z.data_ptr(), codebook.data_ptr(), b, hw, k, d, f, float(commitment_weight), idx.data_ptr(),
                                    q.data_ptr() if q is not None else None, ws.stats.data_ptr(), scalars.data_ptr(),
                                    ws.buf.data_ptr(), ws.nbytes)
    return idx, q, scalars


class VQTrainStepCF(torch.autograd.Function):
    """VQTrainStep for a channels-first caller: z [b, d, hw] in, q_st [b, d, hw] out, differentiable in z.

    forward: one tiled transpose (z -> x [b hw, d], a temporary) + the fused train step writing q channels-first
    (tvq_train_step_qcf; data-parallel: the same kernel exchanges the statistics over NVLink).  With IN_PLACE the
    kernel reads z itself (tvq_train_step_cf: no copy on either side of the VQ) — parity-tested,
    but its 4-byte cp.async tile fill is slower on B200 than transpose + TMA (86 vs 59 us at 1024 x 75 latents), so it
    is not the default.  backward: ONE kernel (tvq_backward_cfx) on the caller's own z, everything channels-first."""

    IN_PLACE = False

    @staticmethod
    def forward(ctx, z, cb, commitment_weight):
        ctx.set_materialize_grads(False)
        b, d, hw = z.shape
        x = z
        ws = cb._workspace(z.device)
        embed = cb._embed_data()
        k = embed.shape[0]
        prev = torch.empty_like(embed) if ctx.needs_input_grad[0] else None
        n = b * hw
        idx = torch.empty(n, dtype=torch.int64, device=z.device)
        q = torch.empty(b, d, hw, dtype=torch.float32, device=z.device)
        scalars = torch.empty(_lib.NUM_SCALARS, dtype=torch.float32, device=z.device)
        commit = torch.empty((), dtype=torch.float32, device=z.device)
        weighted = torch.empty(1, dtype=torch.float32, device=z.device)
        px = cb._peer_exchange(z.device) if cb._ddp_active() else None
        if cb._ddp_active() and px is None:
            raise RuntimeError("internal: the channels-first step needs the peer exchange when data-parallel")
        tail = (px.peers.data_ptr() if px is not None else None, px.rank if px is not None else 0, px.world if px is not None else 1)
        head = (embed.data_ptr(), cb.cluster_size.data_ptr(), cb.embed_avg.data_ptr(), prev.data_ptr() if prev is not None else None)
        outs = (float(commitment_weight), float(cb.decay), float(cb.eps), idx.data_ptr(), q.data_ptr(), scalars.data_ptr(),
                commit.data_ptr(), weighted.data_ptr(), ws.buf.data_ptr(), ws.nbytes)
        _hint_sm_share(cb, z.device)
        deferred = _defer_begin(cb, px)
        if VQTrainStepCF.IN_PLACE:
            _launch("tvq_train_step_cf", z, z.data_ptr(), *head, b, int(hw), k, d, *outs, *tail)
        else:
            xr = transpose12(z)                           # [b, hw, d]: dropped right after the launch
            _launch("tvq_train_step_qcf", z, xr.data_ptr(), *head, n, k, d, *outs, *tail, int(hw))
        if deferred:
            _defer_finish(cb, px, z, deferred, ws)
        ctx.save_for_backward(x, idx, prev)
        ctx.meta = (b, hw, float(commitment_weight))
        ctx.mark_non_differentiable(idx, scalars)
        return q, idx, scalars, commit, weighted

    @staticmethod
    def backward(ctx, g_q, g_idx, g_scalars, g_commit, g_weighted):
        x, idx, prev = ctx.saved_tensors                  # x: the caller's z [b, d, hw]
        b, hw, w = ctx.meta
        d = x.shape[1]
        if g_commit is None and g_weighted is None:
            return (g_q.contiguous() if g_q is not None else None), None, None
        g_q = g_q.contiguous() if g_q is not None else None
        g_commit = g_commit.contiguous() if g_commit is not None else None
        g_weighted = g_weighted.contiguous() if g_weighted is not None else None
        g_z = torch.empty(b, d, hw, dtype=torch.float32, device=x.device)
        ptr = lambda t: t.data_ptr() if t is not None else None
        _launch("tvq_backward_cfx", x, ptr(g_q), ptr(g_commit), ptr(g_weighted), x.data_ptr(), idx.data_ptr(), prev.data_ptr(),
                                          b, hw, prev.shape[0], d, w, g_z.data_ptr())
        return g_z, None, None
