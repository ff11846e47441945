"""ctypes binding of oracle/vq_canon.c (TEST INFRASTRUCTURE ONLY — see that file's header)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libvq_canon.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "vq_canon.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def norms(codebook) -> np.ndarray:
    e = _f32(codebook)
    k, d = e.shape
    out = np.empty(k, np.float32)
    lib().tvq_canon_norms(_p(e), ctypes.c_int(k), ctypes.c_int(d), _p(out))
    return out


def assign(x, codebook, want_margin: bool = False):
    """Nearest code per row under the canonical rule; optionally (best, second) scores."""
    x, e = _f32(x), _f32(codebook)
    n, d = x.shape
    k = e.shape[0]
    idx = np.empty(n, np.int64)
    best = np.empty(n, np.float32) if want_margin else None
    second = np.empty(n, np.float32) if want_margin else None
    lib().tvq_canon_assign(_p(x), _p(e), ctypes.c_int64(n), ctypes.c_int(k), ctypes.c_int(d), _p(idx),
                           _p(best) if want_margin else None, _p(second) if want_margin else None)
    return (idx, best, second) if want_margin else idx


def scores(x, codebook) -> np.ndarray:
    x, e = _f32(x), _f32(codebook)
    n, d = x.shape
    k = e.shape[0]
    out = np.empty((n, k), np.float32)
    lib().tvq_canon_scores(_p(x), _p(e), ctypes.c_int64(n), ctypes.c_int(k), ctypes.c_int(d), _p(out))
    return out


def apply(x, codebook, idx):
    """(q_st, loss_sum, counts, embed_sum) for given indices; sums in fp64."""
    x, e = _f32(x), _f32(codebook)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    n, d = x.shape
    k = e.shape[0]
    q = np.empty((n, d), np.float32)
    loss = ctypes.c_double(0.0)
    counts = np.empty(k, np.float64)
    esum = np.empty((k, d), np.float64)
    lib().tvq_canon_apply(_p(x), _p(e), _p(idx), ctypes.c_int64(n), ctypes.c_int(k), ctypes.c_int(d), _p(q),
                          ctypes.byref(loss), _p(counts), _p(esum))
    return q, loss.value, counts, esum
