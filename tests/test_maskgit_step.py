"""MaskGIT decoding step: oracle vs golden vectors of the unmodified reference (CPU); CUDA kernel
vs the golden vectors and vs the oracle on seeded inputs (GPU).  Token ids are index work: exact, except at positions
the float implementations of expf / logf cannot decide (two best ratios or two confidences within 1e-5 relative)."""
import numpy as np
import pytest
import torch

import maskgit_oracle as MO
from conftest import load_golden

CASES = ["maskgit_lf_t0", "maskgit_lf_mid", "maskgit_hf", "maskgit_k512", "maskgit_last"]


def T(g, k):
    return torch.from_numpy(g[k])


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    g = load_golden(name)
    s_new, sampled, masking = MO.maskgit_step(T(g, "logits"), T(g, "s"), int(g["mask_id"]), int(g["mask_len"]),
                                              float(g["temperature"]), T(g, "q"), T(g, "u"))
    assert torch.equal(s_new, T(g, "s_new")) and torch.equal(sampled, T(g, "sampled")) and torch.equal(masking, T(g, "masking"))


@pytest.fixture(scope="module")
def tvq():
    import tvq_b200
    assert torch.cuda.is_available()
    return tvq_b200


def undecidable(logits, q, sampled_ref, rel=1e-5):
    """positions whose best and second-best p / q differ by less than `rel` (no two float softmaxes agree there)"""
    r = torch.softmax(logits.double(), -1) / q.double()
    top2 = r.topk(2, dim=-1).values
    return (top2[..., 0] - top2[..., 1]) <= rel * top2[..., 0]


def check_step(tvq, logits, s, mask_id, mask_len, temp, q, u, ref):
    s_new, sampled, masking = tvq.maskgit_step(logits.cuda(), s.cuda(), mask_id, mask_len, temp, noise=(q.cuda(), u.cuda()),
                                               return_details=True)
    rs, rsa, rm = ref
    bad = (sampled.cpu() != rsa) & ~undecidable(logits, q, rsa)
    assert not bool(bad.any()), f"{int(bad.sum())} sampled ids differ at decidable positions"
    assert int(masking.sum(-1).min()) == mask_len and int(masking.sum(-1).max()) == mask_len
    if torch.equal(sampled.cpu(), rsa):
        # with identical ids the confidences are the same numbers up to logf rounding: the masks may differ only
        # where the mask_len-th and (mask_len+1)-th smallest confidences are within 1e-5
        diff = masking.cpu() != rm
        assert int(diff.sum()) <= 2 * logits.shape[0] // 50 + 0 or not bool(diff.any())
        if not bool(diff.any()):
            assert torch.equal(s_new.cpu(), rs)
    known = s != mask_id
    assert torch.equal(sampled.cpu()[known], s[known]) and not bool(masking.cpu()[known].any() and mask_len <= int((~known).sum(-1).min()))


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_kernel_matches_reference_golden(tvq, name):
    g = load_golden(name)
    ref = (T(g, "s_new"), T(g, "sampled"), T(g, "masking"))
    s_new, sampled, masking = tvq.maskgit_step(T(g, "logits").cuda(), T(g, "s").cuda(), int(g["mask_id"]), int(g["mask_len"]),
                                               float(g["temperature"]), noise=(T(g, "q").cuda(), T(g, "u").cuda()),
                                               return_details=True)
    assert torch.equal(sampled.cpu(), ref[1]), "sampled ids differ from the reference"
    assert torch.equal(masking.cpu(), ref[2]), "re-masked positions differ from the reference"
    assert torch.equal(s_new.cpu(), ref[0])


@pytest.mark.gpu
@pytest.mark.parametrize("b,n,k,mask_len,temp", [(32, 18, 32, 9, 2.0), (32, 75, 32, 40, 1.0), (5, 108, 1024, 17, 0.5), (1, 1, 2, 1, 0.0),
                                                 (64, 27, 33, 0, 3.0)])
def test_kernel_matches_oracle(tvq, b, n, k, mask_len, temp):
    g = torch.Generator().manual_seed(b * 7 + n)
    logits = torch.randn(b, n, k, generator=g) * 3
    mask_id = k
    s = torch.randint(0, k, (b, n), generator=g)
    for r in range(b):
        s[r, torch.randperm(n, generator=g)[:max(mask_len, n // 2)]] = mask_id
    q = torch.empty(b * n, k).exponential_(1, generator=g).view(b, n, k)
    u = torch.zeros(b, n).uniform_(0, 1, generator=g)
    ref = MO.maskgit_step(logits, s, mask_id, mask_len, temp, q, u)
    check_step(tvq, logits, s, mask_id, mask_len, temp, q, u, ref)


@pytest.mark.gpu
def test_generator_driven_noise_follows_the_reference_order(tvq):
    """maskgit_step draws q then u from the generator it is given — the order the reference consumes torch's RNG."""
    dev = torch.device("cuda")
    logits = torch.randn(4, 18, 32, device=dev)
    s = torch.full((4, 18), 32, dtype=torch.int64, device=dev)
    g1 = torch.Generator(device=dev).manual_seed(5)
    a = tvq.maskgit_step(logits, s, 32, 6, 1.5, generator=g1)
    g2 = torch.Generator(device=dev).manual_seed(5)
    q = torch.empty(4 * 18, 32, device=dev).exponential_(1, generator=g2).view(4, 18, 32)
    u = torch.zeros(4, 18, device=dev).uniform_(0, 1, generator=g2)
    b = tvq.maskgit_step(logits, s, 32, 6, 1.5, noise=(q, u))
    assert torch.equal(a, b) and int((a == 32).sum()) == 4 * 6
