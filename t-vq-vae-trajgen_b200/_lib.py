"""ctypes binding of the C ABI in include/tvq.h (libtvq_b200.so, built in-tree by __graft_entry__.build()).

There is no fallback: if the shared library is missing or the device is not a B200 the
import / first call fails loudly.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtvq_b200.so")

# flags of tvq_forward (include/tvq.h)
F_TRAIN, F_WRITE_Q, F_EXACT, F_NO_UMMA, F_GIVEN_IDX = 1, 2, 4, 8, 16
NUM_SCALARS = 8      # [0] commit, [1] perplexity, [2] weight*commit, [4:6] uint32 diagnostics

EXPORTS = ("tvq_abi_version", "tvq_error_string", "tvq_device_check", "tvq_workspace_bytes", "tvq_forward",
           "tvq_train_step", "tvq_train_step_dp", "tvq_ema_update", "tvq_exchange_bytes", "tvq_ema_update_dp", "tvq_backward", "tvq_gather", "tvq_gather_checked", "tvq_set_peer_timeout", "tvq_hint_max_ctas", "tvq_snake_forward", "tvq_snake_backward", "tvq_hint_defer_exchange", "tvq_ema_finalize_dp", "tvq_neg_dist", "tvq_reseed", "tvq_frontend", "tvq_band_istft", "tvq_band_istft_backward", "tvq_band_istft_frames", "tvq_band_istft_frames_backward", "tvq_maskgit_step", "tvq_transpose", "tvq_forward_qcf", "tvq_train_step_qcf", "tvq_backward_cf",
           "tvq_forward_cf", "tvq_train_step_cf", "tvq_backward_cfx")

_c = ctypes
_vp, _i, _i64, _u, _f, _d, _sz = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_uint, _c.c_float, _c.c_double, _c.c_size_t
_SIGNATURES = {
    "tvq_abi_version": (_i, []),
    "tvq_error_string": (_c.c_char_p, [_i]),
    "tvq_device_check": (_i, [_i, _c.POINTER(_i)]),
    "tvq_workspace_bytes": (_sz, [_i64, _i, _i]),
    "tvq_forward": (_i, [_vp, _vp, _i64, _i, _i, _u, _f, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "tvq_ema_update": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _d, _d, _vp, _sz, _vp]),
    "tvq_exchange_bytes": (_sz, [_i, _i, _i]),
    "tvq_ema_update_dp": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _d, _d, _vp]),
    "tvq_train_step": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _f, _d, _d, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "tvq_train_step_dp": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _f, _d, _d, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _i, _i,
                              _vp]),
    "tvq_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _f, _vp, _vp]),
    "tvq_gather": (_i, [_vp, _vp, _i64, _i64, _i, _i, _i, _vp, _vp]),
    "tvq_gather_checked": (_i, [_vp, _vp, _i64, _i64, _i, _i, _i, _vp, _vp, _vp]),
    "tvq_set_peer_timeout": (_i, [_d]),
    "tvq_hint_max_ctas": (_i, [_i]),
    "tvq_snake_forward": (_i, [_vp, _vp, _i64, _i, _i64, _i, _vp, _vp]),
    "tvq_snake_backward": (_i, [_vp, _vp, _vp, _i64, _i, _i64, _i, _vp, _vp, _vp]),
    "tvq_hint_defer_exchange": (_i, [_i]),
    "tvq_ema_finalize_dp": (_i, [_i, _vp, _sz, _vp, _i, _i, _vp, _vp, _vp, _i, _i, _d, _d, _vp]),
    "tvq_neg_dist": (_i, [_vp, _vp, _i64, _i, _i, _vp, _vp]),
    "tvq_frontend": (_i, [_vp, _i64, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tvq_band_istft": (_i, [_vp, _i64, _i, _i, _i, _i, _vp, _vp]),
    "tvq_band_istft_backward": (_i, [_vp, _i64, _i, _i, _i, _i, _vp, _vp]),
    "tvq_band_istft_frames": (_i, [_vp, _i64, _i, _i, _i, _i, _i, _vp, _vp]),
    "tvq_band_istft_frames_backward": (_i, [_vp, _i64, _i, _i, _i, _i, _i, _vp, _vp]),
    "tvq_maskgit_step": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _i, _i64, _i, _f, _vp, _vp, _vp, _vp]),
    "tvq_transpose": (_i, [_vp, _i64, _i, _i, _vp, _vp]),
    "tvq_forward_qcf": (_i, [_vp, _vp, _i64, _i, _i, _u, _f, _vp, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "tvq_train_step_qcf": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _f, _d, _d, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _i, _i,
                               _i, _vp]),
    "tvq_backward_cf": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _f, _vp, _vp]),
    "tvq_forward_cf": (_i, [_vp, _vp, _i64, _i, _i, _i, _u, _f, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "tvq_train_step_cf": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _f, _d, _d, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _i, _i,
                              _vp]),
    "tvq_backward_cfx": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _f, _vp, _vp]),
    "tvq_reseed": (_i, [_vp, _vp, _vp, _f, _vp, _i64, _i, _i, _vp]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load libtvq_b200.so; raises if it has not been built (no CPU or eager fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: the B200 VQ kernels are not built. Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` at the repo root (needs nvcc, sm_100a).")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if os.environ.get("TVQ_PEER_TIMEOUT_S"):                 # data-parallel wait bound (include/tvq.h)
            check_rc = lib.tvq_set_peer_timeout(float(os.environ["TVQ_PEER_TIMEOUT_S"]))
            if check_rc != 0:
                raise ValueError("TVQ_PEER_TIMEOUT_S must be a number of seconds >= 0 (0 = wait for ever)")
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().tvq_error_string(rc)
        raise RuntimeError(f"{what} failed ({rc}): {msg.decode() if msg else 'unknown'}")
