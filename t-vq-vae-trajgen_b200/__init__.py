"""tvq_b200 — B200-native (sm_100a) vector quantisation for TimeVQVAE.

Public surface = the reference's: `VectorQuantize` (timevqvae/models/vq.py) and the `quantize`
layout glue (timevqvae/utils/train_utils.py), running on hand-written CUDA kernels behind the C
ABI of include/tvq.h.  The directory name carries the project name (`t-vq-vae-trajgen_b200`);
import it as `tvq_b200` through the loader module of that name at the repository root.
"""
from . import _lib
from .functional import (PeerExchange, VQTrainStep, Workspace, stats_len, stats_offset, vq_backward, vq_ema_update,
                         vq_ema_update_dp,
                         vq_forward_raw, vq_gather, vq_neg_dist, vq_reseed, vq_train_step_raw)
from .glue import (band_timefreq_to_time, decode_tokens, lf_hf_frontend, maskgit_step, quantize, time_to_timefreq,
                   timefreq_to_time)
from .vq import EuclideanCodebook, VectorQuantize
from . import stage1
from .stage1 import Stage1, Stage1Trainer

__all__ = ["Stage1", "Stage1Trainer", "stage1", "VectorQuantize", "EuclideanCodebook", "quantize", "decode_tokens", "lf_hf_frontend", "time_to_timefreq", "timefreq_to_time", "band_timefreq_to_time", "maskgit_step", "vq_forward_raw", "vq_ema_update", "vq_ema_update_dp", "PeerExchange",
           "vq_train_step_raw", "vq_backward", "vq_gather", "vq_neg_dist", "vq_reseed", "Workspace", "VQTrainStep", "stats_len",
           "stats_offset", "_lib"]
