"""CPU-side checks of the drop-in boundary: the shared library builds for sm_100a, loads, and
exports exactly the symbols include/tvq.h declares; the host module mirrors the reference's
constructor / state_dict surface and refuses to run without CUDA (no fallback path)."""
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def tvq():
    import __graft_entry__ as g
    g.build()
    import tvq_b200
    return tvq_b200


def header_functions():
    src = open(os.path.join(ROOT, "include", "tvq.h")).read()
    return re.findall(r"TVQ_API\s+[\w\s\*]+?\b(tvq_\w+)\s*\(", src)


def test_library_exports_every_declared_symbol(tvq):
    declared = header_functions()
    assert len(declared) >= 10
    lib = tvq._lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/tvq.h but not exported"
    assert set(declared) == set(tvq._lib.EXPORTS)
    assert lib.tvq_abi_version() == 2
    assert b"unsupported" in lib.tvq_error_string(-1)
    assert lib.tvq_workspace_bytes(0, 32, 128) >= 64 + 32 * 4


def test_library_is_sm100a_with_no_other_arch(tvq):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", tvq._lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_constructor_signature_matches_reference_surface(tvq):
    """Keyword names and defaults of timevqvae/models/vq.py:256-278."""
    sig = inspect.signature(tvq.VectorQuantize.__init__)
    expected = dict(codebook_dim=None, heads=1, decay=0.8, eps=1e-5, kmeans_init=False, kmeans_iters=10,
                    use_cosine_sim=False, threshold_ema_dead_code=0, channel_last=True, accept_image_fmap=False,
                    commitment_weight=1.0, orthogonal_reg_weight=0.0, orthogonal_reg_active_codes_only=False,
                    orthogonal_reg_max_codes=None, sample_codebook_temp=0.0, sync_codebook=False, emb_dropout=0.0)
    for k, v in expected.items():
        assert sig.parameters[k].default == v, k
    assert any(p.kind is inspect.Parameter.VAR_KEYWORD for p in sig.parameters.values())
    # trainers/stage1.py:56-61 passes the whole VQ-VAE config block
    vq = tvq.VectorQuantize(128, 32, n_fft=4, codebook_sizes={"lf": 32, "hf": 32})
    assert vq.codebook_size == 32 and vq.codebook.shape == (32, 128)
    assert isinstance(vq.project_in, torch.nn.Identity) and isinstance(vq.project_out, torch.nn.Identity)


def test_state_dict_keys_and_checkpoint_round_trip(tvq):
    """`_codebook.{initted,cluster_size,embed_avg,embed}` (+ projections when projected)."""
    vq = tvq.VectorQuantize(128, 32)
    assert list(vq.state_dict().keys()) == ["_codebook.initted", "_codebook.cluster_size", "_codebook.embed_avg",
                                            "_codebook.embed"]
    vq2 = tvq.VectorQuantize(128, 32, codebook_dim=64)
    keys = set(vq2.state_dict().keys())
    assert {"project_in.weight", "project_in.bias", "project_out.weight", "project_out.bias"} <= keys
    assert vq2.codebook.shape == (32, 64)
    sd = {k: torch.randn_like(v) for k, v in vq.state_dict().items()}
    vq.load_state_dict(sd)
    assert torch.equal(vq._codebook.embed, sd["_codebook.embed"])
    assert vq._codebook._initted_host is None        # re-read from the loaded flag on the next call


def test_no_cpu_fallback(tvq):
    vq = tvq.VectorQuantize(128, 32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vq(torch.randn(2, 3, 128))
    with pytest.raises(NotImplementedError):
        tvq.VectorQuantize(130, 32)


def test_packed_statistics_layout(tvq):
    assert tvq.stats_offset(32) == 32 and tvq.stats_offset(30) == 32 and tvq.stats_offset(1) == 4
    assert tvq.stats_len(32, 128) == 32 + 32 * 128


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "t-vq-vae-trajgen_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "vq_oracle" not in text and "vq_canon" not in text.replace("oracle/vq_canon.c", ""), f
