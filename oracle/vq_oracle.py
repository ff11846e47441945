"""CPU oracle for the TimeVQVAE vector-quantisation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (`t-vq-vae-trajgen_b200/`)
imports this file; only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may.

This is a functional restatement (explicit state dict, pure functions) of the
reference's algorithm, built from the same six torch primitives the reference
uses so that on one machine it is *bitwise* equal to the reference module:

    reference                                      here
    ---------------------------------------------  -------------------------
    timevqvae/models/vq.py:210-214  (dist)         neg_sq_dist
    timevqvae/models/vq.py:51-56,216-222 (assign)  choose_codes
    timevqvae/models/vq.py:223-225 (one-hot,gather) codebook_step
    timevqvae/models/vq.py:59-64,227-243 (EMA)     ema_update
    timevqvae/models/vq.py:67-75,181-195 (expiry)  sample_rows / expire_codes
    timevqvae/models/vq.py:78-106 (kmeans)         kmeans_init
    timevqvae/models/vq.py:246-249 (perplexity)    perplexity_from_onehot
    timevqvae/models/vq.py:325-407 (wrapper)       vq_forward
    timevqvae/utils/train_utils.py:338-358 (glue)  quantize_glue
    timevqvae/models/maskgit.py:465-470 (decode)   decode_gather

Parity pin: `oracle/gen_golden.py` runs the *unmodified* reference module in
this container and stores inputs/outputs under `tests/golden/`;
`tests/test_oracle_golden.py` checks this restatement against every one of
them (the reference ships no tests or golden vectors of its own; its only
known answer, the `87` at vq.py:421, is part of the fixture set).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional, Tuple

import torch
import torch.nn.functional as F

State = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------- state

def new_state(codebook_size: int, dim: int, *, kmeans_init: bool = False,
              generator: Optional[torch.Generator] = None) -> State:
    """Buffers of EuclideanCodebook.__init__ (vq.py:129-168)."""
    if kmeans_init:
        embed = torch.zeros(codebook_size, dim)
    else:
        embed = torch.randn(codebook_size, dim, generator=generator)
    return {
        "initted": torch.Tensor([not kmeans_init]),
        "cluster_size": torch.zeros(codebook_size),
        "embed_avg": embed.clone(),
        "embed": embed,
    }


def clone_state(state: State) -> State:
    return {k: v.clone() for k, v in state.items()}


# ------------------------------------------------------------------------ distance

def neg_sq_dist(flatten: torch.Tensor, embed_kd: torch.Tensor) -> torch.Tensor:
    """-(|x|^2 - (2x)@e^T + |e|^2), N x K fp32 (vq.py:210-214).

    `2 * flatten @ et` parses as `(2*flatten) @ et`; the codebook enters as the
    transposed *view* exactly as in the reference so that the reduction layouts
    (and hence the bits) agree.
    """
    et = embed_kd.t()
    return -(flatten.pow(2).sum(1, keepdim=True) - 2 * flatten @ et + et.pow(2).sum(0, keepdim=True))


def choose_codes(dist: torch.Tensor, temperature: float) -> torch.Tensor:
    """argmax (first index on ties) or Categorical sample (vq.py:51-56)."""
    if temperature == 0:
        return dist.argmax(dim=-1)
    return torch.distributions.categorical.Categorical(logits=dist / temperature).sample()


def assign_chunked(flatten: torch.Tensor, embed_kd: torch.Tensor, max_dist_bytes: int = 1 << 30
                   ) -> torch.Tensor:
    """Row-chunked deterministic assignment for N*K too large to materialise.

    Identical indices to the un-chunked path (each row's distances depend only
    on that row); used by the CPU baseline for the large sweep points.
    """
    n, k = flatten.shape[0], embed_kd.shape[0]
    rows = max(1, min(n, max_dist_bytes // (4 * k)))
    out = torch.empty(n, dtype=torch.long)
    for s in range(0, n, rows):
        out[s:s + rows] = neg_sq_dist(flatten[s:s + rows], embed_kd).argmax(dim=-1)
    return out


# ----------------------------------------------------------------------------- EMA

def ema_update(state: State, flatten: torch.Tensor, onehot: torch.Tensor, *, decay: float,
               eps: float, all_reduce: Optional[Callable[[torch.Tensor], None]] = None) -> None:
    """In-place EMA codebook update (vq.py:227-242, helpers :59-64)."""
    k = state["embed"].shape[0]
    counts = onehot.sum(0)
    if all_reduce is not None:
        all_reduce(counts)
    state["cluster_size"].mul_(decay).add_(counts, alpha=(1 - decay))
    embed_sum = flatten.t() @ onehot                       # D x K
    if all_reduce is not None:
        all_reduce(embed_sum)
    state["embed_avg"].mul_(decay).add_(embed_sum.t(), alpha=(1 - decay))
    cs = state["cluster_size"]
    smoothed = (cs + eps) / (cs.sum() + k * eps) * cs.sum()
    state["embed"].copy_(state["embed_avg"] / smoothed.unsqueeze(1))


def perplexity_from_onehot(onehot: torch.Tensor) -> torch.Tensor:
    """exp(-sum p log(p + 1e-10)) of the batch code histogram (vq.py:246-247)."""
    p = torch.mean(onehot, dim=0)
    return torch.exp(-torch.sum(p * torch.log(p + 1e-10)))


# -------------------------------------------------------------- expiry and k-means

def sample_rows(samples: torch.Tensor, num: int) -> torch.Tensor:
    """vq.py:67-75 — consumes the global torch RNG exactly like the reference."""
    n = samples.shape[0]
    if n >= num:
        pick = torch.randperm(n, device=samples.device)[:num]
    else:
        pick = torch.randint(0, n, (num,), device=samples.device)
    return samples[pick]


def expire_codes(state: State, batch: torch.Tensor, threshold: float) -> None:
    """Dead-code re-seeding: only `embed` is replaced (vq.py:181-195)."""
    if threshold == 0:
        return
    dead = state["cluster_size"] < threshold
    if not torch.any(dead):
        return
    rows = batch.reshape(-1, batch.shape[-1])
    k = state["embed"].shape[0]
    state["embed"].copy_(torch.where(dead[..., None], sample_rows(rows, k), state["embed"]))


def kmeans_init(samples: torch.Tensor, num_clusters: int, num_iters: int = 10
                ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Lloyd iterations with the reference's direct-difference distance (vq.py:78-106)."""
    dim = samples.shape[-1]
    means = sample_rows(samples, num_clusters)
    bins = None
    for _ in range(num_iters):
        diffs = samples[:, None, :] - means[None, :, :]
        dists = -(diffs ** 2).sum(dim=-1)
        buckets = dists.max(dim=-1).indices
        bins = torch.bincount(buckets, minlength=num_clusters)
        empty = bins == 0
        denom = bins.masked_fill(empty, 1)
        sums = buckets.new_zeros(num_clusters, dim, dtype=samples.dtype)
        sums.scatter_add_(0, buckets[:, None].expand(-1, dim), samples)
        sums = sums / denom[..., None]
        means = torch.where(empty[..., None], means, sums)
    return means, bins


# ------------------------------------------------------------------- codebook step

def codebook_step(state: State, x: torch.Tensor, *, training: bool, decay: float = 0.8,
                  eps: float = 1e-5, svq_temp: Optional[float] = None,
                  threshold_ema_dead_code: float = 0, kmeans_iters: int = 10,
                  all_reduce: Optional[Callable[[torch.Tensor], None]] = None):
    """EuclideanCodebook.forward (vq.py:198-251).

    Returns (quantize, embed_ind, perplexity, onehot).  `quantize` is gathered
    from the codebook *before* this step's EMA update (vq.py:225 precedes :242).
    """
    shape, dtype = x.shape, x.dtype
    flatten = x.reshape(-1, shape[-1])
    k = state["embed"].shape[0]

    if not bool(state["initted"]):                                     # vq.py:171-179
        means, bins = kmeans_init(flatten, k, kmeans_iters)
        state["embed"].copy_(means)
        state["embed_avg"].copy_(means.clone())
        state["cluster_size"].copy_(bins)
        state["initted"].copy_(torch.Tensor([True]))

    dist = neg_sq_dist(flatten, state["embed"])
    temp = 0.0 if not svq_temp else svq_temp                           # vq.py:216
    ind = choose_codes(dist, temp)
    onehot = F.one_hot(ind, k).type(dtype)
    ind = ind.view(*shape[:-1])
    quantize = F.embedding(ind, state["embed"])

    if training:
        ema_update(state, flatten, onehot, decay=decay, eps=eps, all_reduce=all_reduce)
        expire_codes(state, x, threshold_ema_dead_code)

    return quantize, ind, perplexity_from_onehot(onehot).detach(), onehot.detach()


# ------------------------------------------------------------------------- wrapper

def vq_forward(state: State, x: torch.Tensor, *, training: bool, heads: int = 1,
               commitment_weight: float = 1.0, decay: float = 0.8, eps: float = 1e-5,
               svq_temp: Optional[float] = None, threshold_ema_dead_code: float = 0,
               kmeans_iters: int = 10, channel_last: bool = True, accept_image_fmap: bool = False,
               project_in: Optional[Callable] = None, project_out: Optional[Callable] = None,
               all_reduce: Optional[Callable[[torch.Tensor], None]] = None):
    """VectorQuantize.forward (vq.py:325-407) without the orthogonal-reg branch.

    Returns (quantize, embed_ind, vq_loss dict, perplexity) with the reference's
    shapes: `loss` is a shape-[1] tensor, `commit_loss` a 0-d tensor (train) or
    the float 0.0 (eval).
    """
    need_transpose = (not channel_last) and (not accept_image_fmap)
    vq_loss = {"loss": torch.tensor([0.0], requires_grad=training), "commit_loss": 0.0,
               "orthogonal_reg_loss": 0.0}
    if accept_image_fmap:
        b, c, height, width = x.shape
        x = x.permute(0, 2, 3, 1).reshape(b, height * width, c)
    if need_transpose:
        x = x.transpose(1, 2)
    if project_in is not None:
        x = project_in(x)
    if heads > 1:
        b, n, hd = x.shape
        x = x.reshape(b, n, heads, hd // heads).permute(0, 2, 1, 3).reshape(b * heads, n, hd // heads)

    quantize, ind, perplexity, _ = codebook_step(
        state, x, training=training, decay=decay, eps=eps, svq_temp=svq_temp,
        threshold_ema_dead_code=threshold_ema_dead_code, kmeans_iters=kmeans_iters,
        all_reduce=all_reduce)

    if training:
        quantize = x + (quantize - x).detach()                          # vq.py:358-360
        if commitment_weight > 0:
            commit = F.mse_loss(quantize.detach(), x)                   # vq.py:364
            vq_loss["commit_loss"] = commit
            vq_loss["loss"] = vq_loss["loss"] + commit * commitment_weight

    if heads > 1:
        bh, n, d = quantize.shape
        quantize = quantize.reshape(bh // heads, heads, n, d).permute(0, 2, 1, 3).reshape(bh // heads, n, heads * d)
        ind = ind.reshape(bh // heads, heads, n).permute(0, 2, 1)
    if project_out is not None:
        quantize = project_out(quantize)
    if need_transpose:
        quantize = quantize.transpose(1, 2)
    if accept_image_fmap:
        quantize = quantize.reshape(b, height, width, -1).permute(0, 3, 1, 2)
        ind = ind.reshape(b, height, width, *ind.shape[2:])
    return quantize, ind, vq_loss, perplexity


# ------------------------------------------------------------------ boundary glue

def quantize_glue(z: torch.Tensor, step: Callable, transpose_channel_length_axes: bool = False):
    """`quantize()` layout glue (utils/train_utils.py:338-358).

    `step(x_bnd)` is any callable with VectorQuantize.forward's return tuple.
    """
    rank = z.dim() - 2
    if rank == 2:
        b, c, h, w = z.shape
        zq, ind, loss, ppl = step(z.permute(0, 2, 3, 1).reshape(b, h * w, c))
        zq = zq.reshape(b, h, w, c).permute(0, 3, 1, 2)
    elif rank == 1:
        if transpose_channel_length_axes:
            z = z.transpose(1, 2)
        zq, ind, loss, ppl = step(z)
        if transpose_channel_length_axes:
            zq = zq.transpose(1, 2)
    else:
        raise ValueError
    return zq, ind, loss, ppl


def decode_gather(tokens: torch.Tensor, embed_kd: torch.Tensor, h: int, w: int,
                  project_out: Optional[Callable] = None) -> torch.Tensor:
    """Token ids -> decoder input (models/maskgit.py:465-470): (b,n) -> (b,c,h,w)."""
    zq = F.embedding(tokens, embed_kd)
    if project_out is not None:
        zq = project_out(zq)
    b, n, c = zq.shape
    return zq.transpose(1, 2).reshape(b, c, h, w)


# -------------------------------------------------------------- analysis helpers

def top2_margin_ulps(dist: torch.Tensor) -> torch.Tensor:
    """Per-row gap between the best and second-best fp32 score, in ulps of the best.

    Rows with a gap of a few ulps are un-decidable between two fp32
    implementations that sum in different orders (the reference's own CPU and
    CUDA paths included); parity tests report them separately.
    """
    top = torch.topk(dist, 2, dim=-1).values
    best, second = top[:, 0], top[:, 1]
    ulp = torch.abs(torch.nextafter(best, torch.full_like(best, math.inf)) - best)
    return (best - second) / ulp


def vq_backward_formula(g_q: torch.Tensor, g_loss: torch.Tensor, x: torch.Tensor,
                        q_st: torch.Tensor, commitment_weight: float) -> torch.Tensor:
    """Closed form of autograd through vq_forward (SURVEY a-7): ST identity + commit loss."""
    return g_q + g_loss * commitment_weight * (2.0 / x.numel()) * (x - q_st)
