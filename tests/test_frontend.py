"""STFT LF/HF front end: oracle vs the reference's golden vectors (CPU), CUDA kernel vs the
golden vectors, the oracle on seeded inputs, and size-independent properties at the BASELINE batch (GPU).
Tolerance: 1e-5 * max|reference| (fp32 filter-bank arithmetic against torch's FFT-based stft / istft)."""
import numpy as np
import pytest
import torch

import frontend_oracle as FO
from conftest import load_golden

CASES = ["frontend_cfg1", "frontend_nfft8", "frontend_odd", "frontend_interp"]
KEYS = ["xf", "enc_in_l", "enc_in_h", "x_l", "x_h"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    g = load_golden(name)
    out = FO.frontend(g["x"], int(g["n_fft"]))
    for k in KEYS:
        assert out[k].shape == g[k].shape
        np.testing.assert_allclose(out[k], g[k], rtol=0, atol=2e-6 * max(1.0, float(np.abs(g[k]).max())), err_msg=k)


@pytest.fixture(scope="module")
def tvq():
    import tvq_b200
    assert torch.cuda.is_available()
    return tvq_b200


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_kernel_matches_reference_golden(tvq, name):
    g = load_golden(name)
    x = torch.from_numpy(g["x"]).cuda()
    out = tvq.lf_hf_frontend(x, int(g["n_fft"]))
    for k in KEYS:
        ref = torch.from_numpy(g[k])
        torch.testing.assert_close(out[k].cpu(), ref, rtol=0, atol=1e-5 * max(1.0, float(ref.abs().max())), msg=lambda m: f"{k}: {m}")
    xf = tvq.time_to_timefreq(x, int(g["n_fft"]), x.shape[1])
    assert torch.equal(xf, out["xf"])


@pytest.mark.gpu
@pytest.mark.parametrize("b,c,l,n_fft", [(1, 1, 5, 4), (7, 4, 200, 4), (3, 2, 64, 16), (2, 5, 333, 8), (4, 1, 97, 32)])
def test_kernel_matches_oracle(tvq, b, c, l, n_fft):
    torch.manual_seed(b * 100 + l)
    x = torch.rand(b, c, l) * 2 - 1
    ref = FO.frontend(x.numpy(), n_fft)
    out = tvq.lf_hf_frontend(x.cuda(), n_fft)
    for k in KEYS:
        r = torch.from_numpy(ref[k])
        torch.testing.assert_close(out[k].cpu(), r, rtol=0, atol=1e-5 * max(1.0, float(r.abs().max())), msg=lambda m: f"{k}: {m}")
    only = tvq.lf_hf_frontend(x.cuda(), n_fft, want=("x_l",))
    assert list(only) == ["x_l"] and torch.equal(only["x_l"], out["x_l"])


@pytest.mark.gpu
def test_full_batch_properties(tvq):
    """BASELINE batch (1024 trajectories x 4 channels x 200 steps): the two bands add up to the signal (perfect
    reconstruction of the windowed overlap-add), the LF encoder input repeats bin 0, the HF one pastes bin 1."""
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.rand(1024, 4, 200, device="cuda", generator=g) * 2 - 1
    out = tvq.lf_hf_frontend(x, 4)
    torch.testing.assert_close(out["x_l"] + out["x_h"], x, rtol=0, atol=2e-6)
    xf = out["xf"]
    assert torch.equal(out["enc_in_l"], xf[:, :, [0], :].expand_as(xf))
    assert torch.equal(out["enc_in_h"][:, :, 1:], xf[:, :, 1:]) and torch.equal(out["enc_in_h"][:, :, 0], xf[:, :, 1])
    assert float(xf[:, 1::2, 0].abs().max()) == 0.0 and float(xf[:, 1::2, -1].abs().max()) < 1e-6     # DC / Nyquist are real


# ------------------------------------------------------------------ decoder side: pad_func + ISTFT + interpolate

# istft_dec_*: the decoders' real output widths (384 / 400 frames -> 200 samples, down-sampling interpolation) and an
# up-sampling case, from oracle/gen_golden_stage1.py
ISTFT_CASES = ["istft_cfg1", "istft_nfft8", "istft_interp", "istft_dec_lf", "istft_dec_hf", "istft_dec_up"]


@pytest.mark.parametrize("name", ISTFT_CASES)
def test_band_istft_oracle_matches_reference_golden(name):
    g = load_golden(name)
    l = g["g_y"].shape[-1]
    for band in ("all", "lf", "hf"):
        ref = g["y_" + band]
        np.testing.assert_allclose(FO.band_istft(g["u"], int(g["n_fft"]), band, l), ref, rtol=0,
                                   atol=2e-6 * max(1.0, float(np.abs(ref).max())), err_msg=band)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ISTFT_CASES)
def test_band_istft_kernel_matches_reference_golden(tvq, name):
    """Forward and autograd backward against the unmodified reference (models/vq_vae.py:259-262 under torch autograd)."""
    g = load_golden(name)
    n_fft = int(g["n_fft"])
    gy = torch.from_numpy(g["g_y"]).cuda()
    c, l = gy.shape[1], gy.shape[2]
    for band in ("all", "lf", "hf"):
        u = torch.from_numpy(g["u"]).cuda().requires_grad_(True)
        y = tvq.band_timefreq_to_time(u, n_fft, c, band, l)
        (y * gy).sum().backward()
        ry, rg = torch.from_numpy(g["y_" + band]), torch.from_numpy(g["g_u_" + band])
        torch.testing.assert_close(y.detach().cpu(), ry, rtol=0, atol=1e-5 * max(1.0, float(ry.abs().max())), msg=lambda m: f"y {band}: {m}")
        torch.testing.assert_close(u.grad.cpu(), rg, rtol=0, atol=1e-5 * max(1.0, float(rg.abs().max())), msg=lambda m: f"g_u {band}: {m}")
    if l % (n_fft // 4) == 0 and g["u"].shape[3] == l // (n_fft // 4) + 1:     # no interpolation: the plain ISTFT
        u = torch.from_numpy(g["u"]).cuda()
        torch.testing.assert_close(tvq.timefreq_to_time(u, n_fft, c).cpu(), torch.from_numpy(g["y_all"]), rtol=0, atol=1e-5 * 4)


@pytest.mark.gpu
def test_band_istft_is_adjoint_and_inverts_the_front_end(tvq):
    """BASELINE batch: <A u, g> == <u, A^T g> (the backward is the exact adjoint), and ISTFT(STFT(x)) == x."""
    gen = torch.Generator(device="cuda").manual_seed(9)
    x = torch.rand(1024, 4, 200, device="cuda", generator=gen) * 2 - 1
    xf = tvq.lf_hf_frontend(x, 4, want=("xf",))["xf"]
    torch.testing.assert_close(tvq.band_timefreq_to_time(xf, 4, 4, "all", 200), x, rtol=0, atol=2e-6)
    for band in ("all", "lf", "hf"):
        u = torch.randn(64, 8, 3, 201, device="cuda", generator=gen).requires_grad_(True)
        g = torch.randn(64, 4, 200, device="cuda", generator=gen)
        y = tvq.band_timefreq_to_time(u, 4, 4, band, 200)
        (gu,) = torch.autograd.grad(y, u, g)
        lhs, rhs = float((y.detach().double() * g.double()).sum()), float((u.detach().double() * gu.double()).sum())
        assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1.0), (band, lhs, rhs)
