"""Load the UNMODIFIED reference (timevqvae package) for timing / checking.  TEST INFRASTRUCTURE ONLY: imported by tests/,
bench.py's cpu_baseline / --impl reference legs and oracle/gen_golden*.py — never by the product.

The reference is pure Python; `oracle/build_ref.py` copies its package (sources untouched) into the git-ignored
oracle/_ref/ so that it travels to the GPU box, where /root/reference does not exist.  Third-party modules the reference
imports but this image lacks (lightning, mlflow, matplotlib, traffic, x_transformers, ...: SURVEY section 0) are stubbed in
sys.meta_path; none of them is touched by the VQ / stage-1 arithmetic."""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = (os.path.join(HERE, "_ref"), "/root/reference")


class _Any:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Any()

    def __getattr__(self, n):
        return _Any()


class _StubModule(types.ModuleType):
    def __getattr__(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Any


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    ROOTS = ("mlflow", "lightning", "matplotlib", "traffic", "x_transformers", "altair", "cartes", "cartopy", "numba",
             "seaborn", "bluesky", "mpl_toolkits", "openap", "pyproj", "shapely", "geopy")

    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in self.ROOTS:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, m):
        if m.__name__ == "lightning":
            import torch.nn as nn
            m.LightningModule = nn.Module


def reference_root():
    """Directory that holds the `timevqvae` package (oracle/_ref first, then /root/reference), or None."""
    for root in CANDIDATES:
        if os.path.isfile(os.path.join(root, "timevqvae", "models", "vq.py")):
            return root
    return None


_loaded = None


def load():
    """(vq module, stage1 module, train_utils module, root) of the unmodified reference; raises if it is not available."""
    global _loaded
    if _loaded is None:
        root = reference_root()
        if root is None:
            raise ImportError("reference not available: run `python oracle/build_ref.py` where /root/reference exists")
        import warnings
        warnings.filterwarnings("ignore", category=FutureWarning)
        if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
            sys.meta_path.insert(0, _StubFinder())
        if root not in sys.path:
            sys.path.insert(0, root)
        import timevqvae.models.vq as ref_vq
        import timevqvae.trainers.stage1 as ref_stage1
        import timevqvae.utils.train_utils as ref_tu
        _loaded = (ref_vq, ref_stage1, ref_tu, root)
    return _loaded
