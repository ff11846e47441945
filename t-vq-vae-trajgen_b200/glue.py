"""Layout glue around the VQ call sites of the reference.

`quantize`  mirrors timevqvae/utils/train_utils.py:338-358 (b c h w <-> b (h w) c around the VQ),
`decode_tokens` mirrors the token -> decoder-input hand-off of timevqvae/models/maskgit.py:465-470.
"""
from __future__ import annotations

from typing import Union

import torch

from . import functional as TF


def quantize(z, vq_model, transpose_channel_length_axes: bool = False, svq_temp: Union[float, None] = None):
    """z: (b c h w) or (b c l) encoder output -> (z_q, indices, vq_loss, perplexity)."""
    input_dim = z.dim() - 2
    if input_dim == 2:
        b, c, h, w = z.shape
        z = z.permute(0, 2, 3, 1).reshape(b, h * w, c)
        z_q, indices, vq_loss, perplexity = vq_model(z, svq_temp)
        z_q = z_q.reshape(b, h, w, -1).permute(0, 3, 1, 2)
    elif input_dim == 1:
        if transpose_channel_length_axes:
            z = z.transpose(1, 2)
        z_q, indices, vq_loss, perplexity = vq_model(z, svq_temp)
        if transpose_channel_length_axes:
            z_q = z_q.transpose(1, 2)
    else:
        raise ValueError
    return z_q, indices, vq_loss, perplexity


@torch.no_grad()
def decode_tokens(s: torch.Tensor, vq_model, h: int, w: int) -> torch.Tensor:
    """Token ids (b, n) -> decoder input (b, c, h, w): gather + project_out + 'b n c -> b c h w'.

    Without a projection the gather kernel writes the decoder layout directly (one pass instead
    of the reference's gather + two rearranges).
    """
    embed = vq_model._codebook._embed_data()
    s = s.contiguous()
    if isinstance(vq_model.project_out, torch.nn.Identity):
        zq = TF.vq_gather(s, embed, channels_first=True)             # (b, c, n)
    else:
        zq = vq_model.project_out(TF.vq_gather(s, embed)).transpose(1, 2)
    b, c, n = zq.shape
    return zq.reshape(b, c, h, w)
