"""`VectorQuantize` / `EuclideanCodebook` with the reference's module surface, running on the
B200 kernels of libtvq_b200.so.

Drop-in for /root/reference/timevqvae/models/vq.py:124-407: same constructor keywords (unknown
ones are swallowed, as `**config["VQ-VAE"]` needs, trainers/stage1.py:56-61), same call
`vq(x, svq_temp=None) -> (quantize, embed_ind, vq_loss, perplexity)`, same attributes read by the
callers (`_codebook.embed`, `project_out`, `codebook`, `_codebook.perplexity`, `codebook_size`) and the
same four buffers `_codebook.{initted,cluster_size,embed_avg,embed}` so Lightning checkpoints of
the reference load unchanged.

What differs is where the arithmetic runs: distance + argmin + gather + straight-through +
commitment loss + EMA statistics are ONE fused kernel, the EMA normalisation a second one; the
N x K `dist` and one-hot matrices of the reference are never materialised.  Host-side torch ops
remain only for what the reference itself leaves to `nn.Linear` / `rearrange` (projections, head
split, layout) and for the RNG-consuming branches (k-means seeding, dead-code sampling,
`svq_temp` sampling) so that they share torch's RNG stream with the reference.
"""
from __future__ import annotations

import os
import warnings
from typing import Optional, Union

import torch
import torch.distributed as distributed
import torch.nn.functional as F
from torch import nn

from . import functional as TF


def _default(val, d):
    return val if val is not None else d


def sample_vectors(samples: torch.Tensor, num: int) -> torch.Tensor:
    """Row indices for k-means seeding / dead-code replacement; same RNG calls as vq.py:67-75."""
    n, device = samples.shape[0], samples.device
    if n >= num:
        indices = torch.randperm(n, device=device)[:num]
    else:
        indices = torch.randint(0, n, (num,), device=device)
    return indices


def sample_rows(samples: torch.Tensor, num: int, ddp: bool) -> torch.Tensor:
    """`num` rows of `samples` (vq.py:67-75).  Data-parallel: rank 0's draw FROM RANK 0's BATCH is broadcast, so the
    replicas re-seed / initialise with identical vectors — the reference draws rank-locally and its replicas would
    diverge (SURVEY section 8 e, caveats).  Every rank consumes its RNG exactly as the reference does."""
    rows = samples[sample_vectors(samples, num)].contiguous()
    if ddp:
        distributed.broadcast(rows, src=0)
    return rows


def orthogonal_loss_fn(t: torch.Tensor) -> torch.Tensor:
    """Eq. (2) of arXiv:2112.00384 as used at vq.py:112-118 (off by default; plain torch)."""
    n = t.shape[0]
    normed = F.normalize(t, p=2, dim=-1)
    cosine = normed @ normed.t()
    return ((cosine - torch.eye(n, device=t.device)) ** 2).sum() / (n ** 2)


class EuclideanCodebook(nn.Module):
    """State + kernels of one Euclidean codebook (vq.py:124-251)."""

    def __init__(self, dim, codebook_size, kmeans_init=False, kmeans_iters=10, decay=0.8, eps=1e-5,
                 threshold_ema_dead_code=2, use_ddp=False, learnable_codebook=False, sample_codebook_temp=0,
                 emb_dropout=0.0):
        super().__init__()
        if dim % 4 != 0 or dim > 256:
            raise NotImplementedError(
                f"codebook_dim={dim}: the B200 kernels need codebook_dim % 4 == 0 and <= 256 (no fallback path)")
        self.decay = decay
        embed = (torch.randn if not kmeans_init else torch.zeros)(codebook_size, dim)
        self.dim = dim
        self.codebook_size = codebook_size
        self.kmeans_iters = kmeans_iters
        self.eps = eps
        self.threshold_ema_dead_code = threshold_ema_dead_code
        self.sample_codebook_temp = sample_codebook_temp
        self.emb_dropout = emb_dropout
        self.use_ddp = use_ddp

        self.register_buffer("initted", torch.Tensor([not kmeans_init]))
        self.register_buffer("cluster_size", torch.zeros(codebook_size))
        self.register_buffer("embed_avg", embed.clone())
        self.learnable_codebook = learnable_codebook
        if learnable_codebook:
            self.embed = nn.Parameter(embed)
        else:
            self.register_buffer("embed", embed)

        self.perplexity = None
        self._last_ind = None
        # Fraction of the SMs the fused training-step launch of THIS codebook may use (None: all).  Two independent
        # quantisers stepped on two streams (stage 1's LF and HF codebooks) set complementary shares, e.g. in proportion to
        # their latents, so that the two persistent launches are resident together (include/tvq.h: tvq_hint_max_ctas).
        self.sm_share: Optional[float] = None
        # Data-parallel fused step with the exchange OFF the critical path (include/tvq.h: tvq_hint_defer_exchange): the
        # forward kernel only publishes this rank's statistics; the wait for the peers, the rank-ordered sum and the EMA
        # update run in tvq_ema_finalize_dp on a side stream, overlapping whatever follows the forward (decoder, backward).
        # Readers of embed / embed_avg / cluster_size are ordered after it automatically (__getattr__ -> join_pending).
        self.defer_exchange = False            # False | True | 1 | 2 (mode of tvq_hint_defer_exchange; True = 2)
        self.__dict__["_pending"] = None       # torch.cuda.Event recorded after the outstanding finalize kernel, or None
        self._side = None                      # the side stream of the finalize kernels
        self._ws: Optional[TF.Workspace] = None
        self._px = None            # TF.PeerExchange, or False once it is known to be unavailable
        # host mirror of `initted` so the hot path never reads a device flag (the reference syncs
        # on `if self.initted:` every call, vq.py:172)
        self._initted_host: Optional[bool] = not kmeans_init

    _STATE = frozenset(("embed", "embed_avg", "cluster_size"))

    def __getattr__(self, name):
        # buffers live in self._buffers, so every `cb.embed` (module code, MaskGIT's `_codebook.embed`, state_dict users)
        # comes through here: order the reader's stream after an outstanding deferred EMA update first
        if name in EuclideanCodebook._STATE and self.__dict__.get("_pending") is not None:
            self.join_pending()
        return super().__getattr__(name)

    def join_pending(self) -> None:
        """Make the current stream wait for the outstanding tvq_ema_finalize_dp of this codebook (no host synchronisation)."""
        ev = self.__dict__.get("_pending")
        if ev is not None:
            self.__dict__["_pending"] = None
            torch.cuda.current_stream(self._buffers["embed_avg"].device).wait_event(ev)

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        self.join_pending()                    # (nn.Module reads self._buffers directly here)
        return super()._save_to_state_dict(destination, prefix, keep_vars)

    def _apply(self, fn, *args, **kwargs):
        self.join_pending()
        return super()._apply(fn, *args, **kwargs)

    # -- plumbing ---------------------------------------------------------------------------
    def _workspace(self, device: torch.device) -> TF.Workspace:
        if self._ws is None or not self._ws.matches(self.codebook_size, self.dim, device):
            self._ws = TF.Workspace(self.codebook_size, self.dim, device)
        return self._ws

    def _ddp_active(self) -> bool:
        return bool(self.use_ddp and distributed.is_available() and distributed.is_initialized()
                    and distributed.get_world_size() > 1)

    def _all_reduce_stats(self, stats: torch.Tensor) -> None:
        """The reference's two all_reduce hooks (vq.py:229, :234) as ONE call on the packed buffer."""
        if self._ddp_active():
            distributed.all_reduce(stats)

    def setup_data_parallel(self, device: torch.device):
        """COLLECTIVE: set up the exchange buffers of the fused NVLink all-reduce + EMA kernels, and agree across the ranks
        on whether they are used.  Every rank of the process group must call it at the same point (the module does, on its
        first data-parallel training step); the outcome is the same on every rank — a rank where symmetric memory is not
        available, or that sets TVQ_NO_PEER_EXCHANGE, takes ALL ranks to the all_reduce + EMA-kernel path, so no rank ever
        spins in the peer kernel while another sits in NCCL.  Returns the PeerExchange or None."""
        if self._px is not None and (self._px is False or self._px.matches(self.codebook_size, self.dim, device)):
            return self._px or None
        def agreed(flag: bool) -> bool:                                      # MIN over the ranks; every rank calls it
            t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device)
            distributed.all_reduce(t, op=distributed.ReduceOp.MIN)
            return int(t.item()) == 1

        px, why = None, ""
        # (1) the local preconditions, agreed on BEFORE anyone enters the rendezvous (itself a collective)
        if distributed.get_backend() != "nccl":
            why = "peer exchange needs CUDA peers (nccl process group)"
        elif os.environ.get("TVQ_NO_PEER_EXCHANGE"):
            why = "TVQ_NO_PEER_EXCHANGE is set"
        elif TF.stats_len(self.codebook_size, self.dim) > (1 << 16):
            why = "statistics too large for the one-shot exchange"
        go = agreed(not why) if distributed.get_backend() == "nccl" else False
        # (2) allocation + rendezvous, then agree on the outcome
        if go:
            try:
                px = TF.PeerExchange(self.codebook_size, self.dim, device)
            except Exception as exc:
                why = str(exc)
            if not agreed(px is not None):
                px = None
        if px is None and distributed.get_rank() == 0:
            warnings.warn("tvq_b200: NVLink peer exchange not used" + (f" ({why})" if why else " (unavailable on a peer)")
                          + "; all ranks use all_reduce + EMA kernel")
        self._px = px if px is not None else False
        return px

    def _peer_exchange(self, device: torch.device):
        """The exchange buffers, or None (then: all_reduce + EMA kernel).  The first call is the collective setup."""
        if self._px is None or (self._px is not False and not self._px.matches(self.codebook_size, self.dim, device)):
            return self.setup_data_parallel(device)
        return self._px or None

    def check_peer_errors(self) -> None:
        """Raise if a data-parallel kernel of this codebook ever gave up waiting for a peer (include/tvq.h:
        tvq_set_peer_timeout).  Reads one word from the device (a synchronisation): call it off the hot path."""
        if self._px:
            self._px.check()

    def _sync_and_update(self, ws, prev) -> None:
        """EMA update from this call's statistics (ws.stats); data-parallel: summed over the ranks first."""
        embed = self._embed_data()
        if self._ddp_active():
            px = self._peer_exchange(embed.device)
            if px is not None:
                TF.vq_ema_update_dp(ws.stats, px, self.cluster_size, self.embed_avg, embed, prev, self.decay, self.eps)
                return
            distributed.all_reduce(ws.stats)
        TF.vq_ema_update(ws.stats, self.cluster_size, self.embed_avg, embed, prev, self.decay, self.eps, ws)

    def _embed_data(self) -> torch.Tensor:
        e = self.embed.data if self.learnable_codebook else self.embed
        if not e.is_contiguous():
            raise ValueError("codebook buffer must be contiguous")
        return e

    def _load_from_state_dict(self, *args, **kwargs):
        self._initted_host = None
        return super()._load_from_state_dict(*args, **kwargs)

    @property
    def embed_onehot(self):
        """N x K one-hot of the last call (vq.py:248). Nothing reads it; materialised on demand only."""
        if self._last_ind is None:
            return None
        return F.one_hot(self._last_ind.reshape(-1), self.codebook_size).type(torch.float32)

    # -- optional branches that consume torch's RNG ---------------------------------------------
    @torch.no_grad()
    def init_embed_(self, data: torch.Tensor, seed_rows: Optional[torch.Tensor] = None) -> None:
        """k-means initialisation (vq.py:171-179, :78-106): Lloyd iterations on the kernels (assign + per-code sums in one
        launch per iteration; the reference's (N, K, D) broadcast difference is never built).

        seed_rows (k int64 row indices into data): the initial means, instead of a draw from torch's generator — so a
        caller (or a parity test: tests/golden/kmeans_init_train.npz) can fix them.
        Data-parallel (sync_codebook on an initialised group): rank 0's seed vectors are broadcast and the per-iteration
        statistics are summed over the ranks — one global k-means, identical replicas (the reference would run R
        independent k-means on the local batches and diverge).
        Deviation: codes are assigned by the canonical three-term rule of DESIGN section 4, the reference's k-means uses
        the direct difference sum((x - m)^2) (vq.py:87-90); they can differ only on a row whose two nearest means are
        within a few ulps — the parity test counts such rows."""
        if self._initted_host is None:
            self._initted_host = bool(self.initted.item())     # one sync, then cached
        if self._initted_host:
            return
        k = self.codebook_size
        ddp = self._ddp_active()
        ws = self._workspace(data.device)
        if seed_rows is not None:
            means = data[seed_rows.to(data.device)].contiguous()
            if ddp:
                distributed.broadcast(means, src=0)
        else:
            means = sample_rows(data, k, ddp)
        off = TF.stats_offset(k)
        bins = None
        for _ in range(self.kmeans_iters):
            TF.vq_forward_raw(data, means, ws, train=True, write_q=False)      # counts + per-code sums
            if ddp:
                distributed.all_reduce(ws.stats)
            bins = ws.stats[:k].clone()
            sums = ws.stats[off:off + k * self.dim].view(k, self.dim)
            empty = bins == 0
            new_means = sums / bins.masked_fill(empty, 1)[:, None]
            means = torch.where(empty[:, None], means, new_means).contiguous()
        self._embed_data().copy_(means)
        self.embed_avg.copy_(means)
        self.cluster_size.copy_(bins)
        self.initted.fill_(1.0)
        self._initted_host = True

    @torch.no_grad()
    def expire_codes_(self, batch_samples: torch.Tensor) -> None:
        """Dead-code re-seed (vq.py:187-195); only `embed` is replaced, as in the reference.  Data-parallel: cluster_size
        is identical on every rank, so all ranks take the branch together and re-seed with rank 0's rows."""
        if self.threshold_ema_dead_code == 0:
            return
        expired = self.cluster_size < self.threshold_ema_dead_code
        if not torch.any(expired):                       # same sync (and same RNG use) as the reference
            return
        flat = batch_samples.reshape(-1, batch_samples.shape[-1]).contiguous()
        k = self.codebook_size
        if self._ddp_active():
            rows = sample_rows(flat, k, True)            # [k, d] vectors, rank 0's
            TF.vq_reseed(rows, torch.arange(k, device=rows.device), self.cluster_size, self.threshold_ema_dead_code,
                         self._embed_data())
            return
        TF.vq_reseed(flat, sample_vectors(flat, k), self.cluster_size, self.threshold_ema_dead_code, self._embed_data())

    # -- the step -----------------------------------------------------------------------------
    def _assign(self, flat: torch.Tensor, svq_temp):
        """Codes chosen by the caller-visible rule of vq.py:216-222; None means 'let the kernel argmin'."""
        temp = 0.0 if not svq_temp else svq_temp
        embed = self._embed_data()
        scoring = embed
        if self.emb_dropout and self.training:
            scoring = F.dropout(embed.t(), self.emb_dropout).t().contiguous()   # vq.py:207-208
        if temp == 0 and scoring is embed:
            return None
        if temp == 0:
            ws = self._workspace(flat.device)
            idx, _, _ = TF.vq_forward_raw(flat, scoring, ws, train=False, write_q=False)
            return idx
        dist = TF.vq_neg_dist(flat, scoring)
        return torch.distributions.categorical.Categorical(logits=dist / temp).sample()    # vq.py:55-56

    def step(self, x: torch.Tensor, svq_temp=None, *, commitment_weight: float = 0.0, straight_through: bool = False):
        """Fused codebook step on x (..., d).

        Returns (quantize, embed_ind, commit_loss or None, weighted_loss[1] or None).  With
        `straight_through` (training VectorQuantize path) quantize is x + (e[idx] - x), commit_loss
        the mean squared difference and weighted_loss = commitment_weight * commit_loss, all
        differentiable in x; otherwise quantize is the plain gather the reference's
        EuclideanCodebook.forward returns.
        """
        shape = x.shape
        flat = x.reshape(-1, shape[-1])
        if flat.dtype != torch.float32:
            flat = flat.float()                              # @autocast(enabled=False) region, vq.py:197
        if not flat.is_contiguous():
            flat = flat.contiguous()
        TF._need(flat, "x")
        if self._initted_host is not True:
            self.init_embed_(flat.detach())
        given = self._assign(flat.detach(), svq_temp)
        ws = self._workspace(flat.device)
        commit = weighted = None
        if self.training:
            if straight_through:
                q, idx, scalars, commit, weighted = TF.VQTrainStep.apply(flat, self, commitment_weight, given)
            else:
                idx, _, scalars = TF.vq_forward_raw(flat.detach(), self._embed_data(), ws, train=True, write_q=False,
                                                    idx=given)
                q = None
                prev = torch.empty_like(self._embed_data())
                self._sync_and_update(ws, prev)
                q = TF.vq_gather(idx.view(1, -1), prev).view(flat.shape)
            self.expire_codes_(x.detach())
        else:
            idx, q, scalars = TF.vq_forward_raw(flat.detach(), self._embed_data(), ws, train=False, write_q=True,
                                                idx=given)
        self.perplexity = scalars[1].detach()
        ind = idx.view(*shape[:-1])
        self._last_ind = ind
        return q.view(shape), ind, commit, weighted

    def forward(self, x, svq_temp: Union[float, None] = None):
        """(quantize, embed_ind) exactly as EuclideanCodebook.forward (vq.py:198-251)."""
        q, ind, _, _ = self.step(x, svq_temp)
        return q, ind


class VectorQuantize(nn.Module):
    """Reference-compatible wrapper (vq.py:255-407)."""

    def __init__(self, dim, codebook_size, codebook_dim=None, heads=1, decay=0.8, eps=1e-5, kmeans_init=False,
                 kmeans_iters=10, use_cosine_sim=False, threshold_ema_dead_code=0, channel_last=True,
                 accept_image_fmap=False, commitment_weight=1.0, orthogonal_reg_weight=0.0,
                 orthogonal_reg_active_codes_only=False, orthogonal_reg_max_codes=None, sample_codebook_temp=0.0,
                 sync_codebook=False, emb_dropout=0.0, defer_exchange=False, **kwargs):
        super().__init__()
        self.heads = heads
        codebook_dim = _default(codebook_dim, dim)
        codebook_input_dim = codebook_dim * heads
        requires_projection = codebook_input_dim != dim
        self.project_in = nn.Linear(dim, codebook_input_dim) if requires_projection else nn.Identity()
        self.project_out = nn.Linear(codebook_input_dim, dim) if requires_projection else nn.Identity()

        self.eps = eps
        self.commitment_weight = commitment_weight
        has_codebook_orthogonal_loss = orthogonal_reg_weight > 0
        self.orthogonal_reg_weight = orthogonal_reg_weight
        self.orthogonal_reg_active_codes_only = orthogonal_reg_active_codes_only
        self.orthogonal_reg_max_codes = orthogonal_reg_max_codes

        # `use_cosine_sim` is accepted and ignored, as in the reference (only the Euclidean class exists, vq.py:300)
        self._codebook = EuclideanCodebook(
            dim=codebook_dim, codebook_size=codebook_size, kmeans_init=kmeans_init, kmeans_iters=kmeans_iters,
            decay=decay, eps=eps, threshold_ema_dead_code=threshold_ema_dead_code, use_ddp=sync_codebook,
            learnable_codebook=has_codebook_orthogonal_loss, sample_codebook_temp=sample_codebook_temp,
            emb_dropout=emb_dropout)
        # (not a reference keyword) data-parallel only: take the statistics exchange off the critical path, see
        # EuclideanCodebook.defer_exchange; a stage-1 config passes it under "VQ-VAE" next to sync_codebook
        self._codebook.defer_exchange = defer_exchange       # False | True (= 2) | 1 | 2: include/tvq.h, tvq_hint_defer_exchange
        self.codebook_size = codebook_size
        self.accept_image_fmap = accept_image_fmap
        self.channel_last = channel_last

    @property
    def codebook(self):
        return self._codebook.embed

    def forward(self, x, svq_temp: Union[float, None] = None):
        """x: (B, N, D) -> (quantize, embed_ind, vq_loss dict, perplexity)   (vq.py:325-407)"""
        device, heads = x.device, self.heads
        need_transpose = not self.channel_last and not self.accept_image_fmap
        vq_loss = {"loss": None, "commit_loss": 0.0, "orthogonal_reg_loss": 0.0}

        if self.accept_image_fmap:
            b, c, height, width = x.shape
            x = x.permute(0, 2, 3, 1).reshape(b, height * width, c)
        if need_transpose:
            x = x.transpose(1, 2)
        x = self.project_in(x)
        if heads > 1:
            b, n, hd = x.shape
            x = x.reshape(b, n, heads, hd // heads).permute(0, 2, 1, 3).reshape(b * heads, n, hd // heads)

        want_commit = self.training and self.commitment_weight > 0
        quantize, embed_ind, commit_loss, weighted = self._codebook.step(
            x, svq_temp, commitment_weight=self.commitment_weight if want_commit else 0.0,
            straight_through=self.training)

        if self.training:
            if want_commit:
                # [0.] + commit * w of vq.py:338,366 — the product comes out of the kernel (scalars[2])
                vq_loss["commit_loss"] = commit_loss
                vq_loss["loss"] = weighted
            if self.orthogonal_reg_weight > 0:
                codebook = self.codebook
                if self.orthogonal_reg_active_codes_only:
                    codebook = codebook[torch.unique(embed_ind)]
                num_codes = codebook.shape[0]
                if self.orthogonal_reg_max_codes is not None and num_codes > self.orthogonal_reg_max_codes:
                    rand_ids = torch.randperm(num_codes, device=device)[: self.orthogonal_reg_max_codes]
                    codebook = codebook[rand_ids]
                orthogonal_reg_loss = orthogonal_loss_fn(codebook)
                vq_loss["orthogonal_reg_loss"] = orthogonal_reg_loss
                base = vq_loss["loss"] if vq_loss["loss"] is not None else torch.zeros(1, device=device)
                vq_loss["loss"] = base + orthogonal_reg_loss * self.orthogonal_reg_weight

        if vq_loss["loss"] is None:
            vq_loss["loss"] = torch.zeros(1, device=device, requires_grad=self.training)
        if heads > 1:
            bh, n, d = quantize.shape
            quantize = quantize.reshape(bh // heads, heads, n, d).permute(0, 2, 1, 3).reshape(bh // heads, n, heads * d)
            embed_ind = embed_ind.reshape(bh // heads, heads, n).permute(0, 2, 1)
        quantize = self.project_out(quantize)
        if need_transpose:
            quantize = quantize.transpose(1, 2)
        if self.accept_image_fmap:
            quantize = quantize.reshape(b, height, width, -1).permute(0, 3, 1, 2)
            embed_ind = embed_ind.reshape(b, height, width, *embed_ind.shape[2:])
        return quantize, embed_ind, vq_loss, self._codebook.perplexity

    def _channels_first_ok(self, z: torch.Tensor, svq_temp) -> bool:
        """Can forward_channels_first take this call?  (plain module, fp32 CUDA, codebook on the resident-codebook kernel)"""
        cb = self._codebook
        return (isinstance(z, torch.Tensor) and z.is_cuda and z.dtype == torch.float32 and z.dim() in (3, 4)
                and not svq_temp and self.heads == 1 and self.channel_last and not self.accept_image_fmap
                and isinstance(self.project_in, nn.Identity) and isinstance(self.project_out, nn.Identity)
                and self.orthogonal_reg_weight == 0 and not cb.learnable_codebook and not cb.emb_dropout
                and cb.threshold_ema_dead_code == 0 and cb._initted_host is True and cb.dim <= 128
                and cb.codebook_size <= (32 if self.training else 64) and z.shape[1] == cb.dim
                and z.numel() > 0 and z.shape[0] * z[0, 0].numel() < 2 ** 31 - 64
                # tvq_backward_cfx keeps the codebook (k x (d+1)) and one index row (hw) in <= 100 KB of shared memory
                and (not self.training or (cb.codebook_size * (cb.dim + 1) + z[0, 0].numel()) * 4 <= 100 * 1024)
                # data-parallel: only with the peer exchange already agreed on (setup_data_parallel: a collective the
                # caller runs first — quantize() does — never a side effect of this predicate)
                and (not (self.training and cb._ddp_active()) or bool(cb._px)))

    def forward_channels_first(self, z: torch.Tensor):
        """z (b, c, h, w) or (b, c, l) -> (z_q in the same layout, embed_ind (b, h*w), vq_loss, perplexity): what
        quantize() (utils/train_utils.py:338-358) computes with two rearranges around forward().  One tiled transpose on
        the way in (or none: TF.VQTrainStepCF.IN_PLACE), the kernel writes z_q channels-first, and the backward is one
        kernel on z itself."""
        cb = self._codebook
        shape = z.shape
        b, c = shape[0], shape[1]
        zz = z.contiguous().view(b, c, -1)
        hw = zz.shape[2]
        vq_loss = {"loss": None, "commit_loss": 0.0, "orthogonal_reg_loss": 0.0}
        if self.training:
            want_commit = self.commitment_weight > 0
            q, idx, scalars, commit, weighted = TF.VQTrainStepCF.apply(zz, cb, self.commitment_weight if want_commit else 0.0)
            if want_commit:
                vq_loss["commit_loss"] = commit
                vq_loss["loss"] = weighted
        else:
            with torch.no_grad():
                if TF.VQTrainStepCF.IN_PLACE:
                    idx, q, scalars = TF.vq_forward_cf(zz.detach(), cb._embed_data(), cb._workspace(z.device), train=False)
                else:
                    x = TF.transpose12(zz.detach()).view(b * hw, c)
                    idx, q, scalars = TF.vq_forward_qcf(x, cb._embed_data(), cb._workspace(z.device), hw, train=False)
        if vq_loss["loss"] is None:
            vq_loss["loss"] = torch.zeros(1, device=z.device, requires_grad=self.training)
        cb.perplexity = scalars[1].detach()
        cb._last_ind = idx.view(b, hw)
        return q.view(shape), idx.view(b, hw), vq_loss, cb.perplexity

    @torch.no_grad()
    def tokenize(self, x: torch.Tensor) -> torch.Tensor:
        """Indices only (what stage 2/3 keep of an eval call, models/maskgit.py:130-134): skips the q write."""
        cb = self._codebook
        x = self.project_in(x)
        shape = x.shape
        flat = x.reshape(-1, shape[-1]).float().contiguous()
        idx, _, scalars = TF.vq_forward_raw(flat, cb._embed_data(), cb._workspace(flat.device), train=False,
                                            write_q=False)
        cb.perplexity = scalars[1]
        return idx.view(*shape[:-1])
