"""Pin the CPU oracle (oracle/vq_oracle.py) against golden vectors produced by the
unmodified reference (oracle/gen_golden.py, run in the build container).

Index/count outputs must be exact.  Float outputs are compared at 1e-6 relative
(they are bitwise equal on the machine that generated them; a different host CPU
may pick a different MKL sgemm kernel and move the last bit).
"""
import hashlib

import numpy as np
import pytest
import torch

import vq_oracle as O
from conftest import load_golden

RT = dict(rtol=1e-6, atol=1e-7)


def T(a):
    return torch.from_numpy(np.asarray(a))


def sha(t):
    return hashlib.sha256(t.detach().contiguous().numpy().tobytes()).hexdigest()


def state_from(g, prefix):
    return {k: T(g[prefix + k]).clone() for k in ("initted", "cluster_size", "embed_avg", "embed")}


def check_state(state, g, prefix):
    for k in ("cluster_size", "embed_avg", "embed"):
        torch.testing.assert_close(state[k], T(g[prefix + k]), **RT, msg=lambda m, k=k: f"{prefix}{k}: {m}")
    assert bool(state["initted"]) == bool(g[prefix + "initted"][0])


def check_loss(vq_loss, g, prefix="out_"):
    torch.testing.assert_close(vq_loss["loss"].detach(), T(g[prefix + "loss"]), **RT)
    assert tuple(vq_loss["loss"].shape) == (1,)
    if bool(g[prefix + "commit_is_tensor"]):
        assert torch.is_tensor(vq_loss["commit_loss"]) and vq_loss["commit_loss"].dim() == 0
        torch.testing.assert_close(vq_loss["commit_loss"].detach(), T(g[prefix + "commit_loss"]), **RT)
    else:
        assert vq_loss["commit_loss"] == 0.0


def test_known_answer_block_vq_py_410_424():
    """The reference's only known answer: vq.py:421 says code 87 comes first."""
    g = load_golden("smoke_main")
    torch.manual_seed(0)
    x = torch.rand((1024, 32, 128))
    state = O.new_state(512, 128)
    if sha(x) != str(g["x_sha"]) or sha(state["embed"]) != str(g["embed_sha"]):
        pytest.skip("torch CPU RNG stream differs from the one that generated the fixture")
    q, ind, loss, ppl = O.vq_forward(state, x, training=True)
    assert ind[0, 0].item() == 87
    assert np.array_equal(ind.numpy().astype(np.int16), g["out_ind"])
    torch.testing.assert_close(q[:2].detach(), T(g["out_q_first"]), **RT)
    torch.testing.assert_close(ppl, T(g["out_perplexity"]), **RT)
    check_loss(loss, g)
    for k in ("cluster_size", "embed_avg", "embed"):
        torch.testing.assert_close(state[k], T(g["post_" + k]), **RT)


@pytest.mark.parametrize("tag", ["lf", "hf"])
def test_config1_stage1_latents_through_glue(tag):
    """configs/config.yaml shapes: encoder output -> quantize() glue -> VQ, train + eval + grad."""
    g = load_golden(f"cfg1_{tag}")
    z = T(g["z"]).clone().requires_grad_(True)
    state = state_from(g, "pre_")
    step = int(g["keep_step"])
    zq, ind, loss, ppl = O.quantize_glue(z, lambda x: O.vq_forward(state, x, training=True))
    assert np.array_equal(ind.numpy().astype(np.int16), g["out_ind"])
    torch.testing.assert_close(zq[::step].detach(), T(g["out_zq_kept"]), **RT)
    torch.testing.assert_close(ppl, T(g["out_perplexity"]), **RT)
    check_loss(loss, g)
    check_state(state, g, "post_")
    gq = torch.randn(zq.shape, generator=torch.Generator().manual_seed(int(g["g_zq_seed"])))
    ((zq * gq).sum() + loss["loss"].sum()).backward()
    torch.testing.assert_close(z.grad[::step], T(g["out_grad_z_kept"]), **RT)
    # eval (tokenise) leaves the buffers alone
    with torch.no_grad():
        zq_e, ind_e, loss_e, ppl_e = O.quantize_glue(z.detach(), lambda x: O.vq_forward(state, x, training=False))
    assert np.array_equal(ind_e.numpy().astype(np.int16), g["eval_ind"])
    torch.testing.assert_close(zq_e[::step], T(g["eval_zq_kept"]), **RT)
    torch.testing.assert_close(ppl_e, T(g["eval_perplexity"]), **RT)
    check_loss(loss_e, g, "eval_")
    check_state(state, g, "post_")


def test_backward_closed_form_matches_autograd():
    g = load_golden("cfg1_lf")
    z = T(g["z"])
    b, c, h, w = z.shape
    x = z.permute(0, 2, 3, 1).reshape(b, h * w, c).clone().requires_grad_(True)
    state = state_from(g, "pre_")
    q, ind, loss, _ = O.vq_forward(state, x, training=True, commitment_weight=0.7)
    gq = torch.randn(q.shape, generator=torch.Generator().manual_seed(3))
    ((q * gq).sum() + 1.3 * loss["loss"].sum()).backward()
    formula = O.vq_backward_formula(gq, torch.tensor(1.3), x.detach(), q.detach(), 0.7)
    torch.testing.assert_close(x.grad, formula, rtol=1e-5, atol=1e-7)


def test_three_training_steps_ema_sequencing():
    g = load_golden("train_3steps")
    state = state_from(g, "pre_")
    w = float(g["commitment_weight"])
    for s in range(3):
        pre_embed = state["embed"].clone()
        q, ind, loss, ppl = O.vq_forward(state, T(g[f"x{s}"]), training=True, commitment_weight=w)
        assert np.array_equal(ind.numpy().astype(np.int16), g[f"out{s}_ind"])
        torch.testing.assert_close(q, T(g[f"out{s}_q"]), **RT)
        torch.testing.assert_close(ppl, T(g[f"out{s}_perplexity"]), **RT)
        check_loss(loss, g, f"out{s}_")
        check_state(state, g, f"post{s}_")
        # quantize comes from the codebook *before* this step's update (vq.py:225 vs :242)
        x = T(g[f"x{s}"])
        assert torch.equal(q, x + (pre_embed[ind] - x))


def test_heads_projection_and_layout_variants():
    g = load_golden("heads2_train")
    state = state_from(g, "pre_")
    q, ind, loss, ppl = O.vq_forward(state, T(g["x"]), training=True, heads=2)
    assert ind.shape == (3, 20, 2) and np.array_equal(ind.numpy().astype(np.int16), g["out_ind"])
    torch.testing.assert_close(q, T(g["out_q"]), **RT)
    check_loss(loss, g)
    check_state(state, g, "post_")

    g = load_golden("proj64_train")
    state = state_from(g, "pre_")
    lin_in, lin_out = torch.nn.Linear(128, 64), torch.nn.Linear(64, 128)
    with torch.no_grad():
        lin_in.weight.copy_(T(g["w_in"])); lin_in.bias.copy_(T(g["b_in"]))
        lin_out.weight.copy_(T(g["w_out"])); lin_out.bias.copy_(T(g["b_out"]))
    x = T(g["x"]).clone().requires_grad_(True)
    q, ind, loss, ppl = O.vq_forward(state, x, training=True, project_in=lin_in, project_out=lin_out)
    assert np.array_equal(ind.numpy().astype(np.int16), g["out_ind"])
    torch.testing.assert_close(q.detach(), T(g["out_q"]), **RT)
    ((q * T(g["g_q"])).sum() + loss["loss"].sum()).backward()
    torch.testing.assert_close(x.grad, T(g["out_grad_x"]), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(lin_in.weight.grad, T(g["grad_w_in"]), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(lin_out.weight.grad, T(g["grad_w_out"]), rtol=1e-5, atol=1e-6)
    check_state(state, g, "post_")

    g = load_golden("image_fmap_train")
    state = state_from(g, "pre_")
    q, ind, loss, ppl = O.vq_forward(state, T(g["x"]), training=True, accept_image_fmap=True)
    assert ind.shape == (2, 3, 5) and np.array_equal(ind.numpy().astype(np.int16), g["out_ind"])
    torch.testing.assert_close(q, T(g["out_q"]), **RT)
    check_state(state, g, "post_")

    g = load_golden("channel_first_train")
    state = state_from(g, "pre_")
    q, ind, loss, ppl = O.vq_forward(state, T(g["x"]), training=True, channel_last=False)
    assert q.shape == (2, 32, 17) and np.array_equal(ind.numpy().astype(np.int16), g["out_ind"])
    torch.testing.assert_close(q, T(g["out_q"]), **RT)
    check_state(state, g, "post_")


@pytest.mark.parametrize("name", ["dead_code_randperm", "dead_code_randint"])
def test_dead_code_reseed_only_replaces_embed(name):
    g = load_golden(name)
    state = state_from(g, "pre_")
    torch.manual_seed(int(g["rng_seed"]))
    q, ind, loss, ppl = O.vq_forward(state, T(g["x"]), training=True, threshold_ema_dead_code=2)
    assert np.array_equal(ind.numpy().astype(np.int16), g["out_ind"])
    torch.testing.assert_close(q, T(g["out_q"]), **RT)
    check_state(state, g, "post_")


def test_kmeans_init():
    g = load_golden("kmeans_init_train")
    state = state_from(g, "pre_")
    assert not bool(state["initted"])
    torch.manual_seed(int(g["rng_seed"]))
    q, ind, loss, ppl = O.vq_forward(state, T(g["x"]), training=True)
    assert np.array_equal(ind.numpy().astype(np.int16), g["out_ind"])
    torch.testing.assert_close(q, T(g["out_q"]), **RT)
    check_state(state, g, "post_")


def test_stochastic_sampling_branch():
    g = load_golden("svq_temp_eval")
    state = state_from(g, "pre_")
    torch.manual_seed(int(g["rng_seed"]))
    q, ind, loss, ppl = O.vq_forward(state, T(g["x"]), training=False, svq_temp=float(g["svq_temp"]))
    assert np.array_equal(ind.numpy().astype(np.int16), g["out_ind"])
    torch.testing.assert_close(q, T(g["out_q"]), **RT)
    torch.testing.assert_close(ppl, T(g["out_perplexity"]), **RT)


def test_decode_gather_layout():
    g = load_golden("decode_gather")
    out = O.decode_gather(T(g["tokens"]).long(), T(g["embed"]), int(g["h"]), int(g["w"]))
    assert torch.equal(out, T(g["out_zq"]))


def test_sync_codebook_two_ranks_equals_summed_statistics():
    """vq.py:229,234 all_reduce hooks: emulate both ranks in one process with a summing hook."""
    g = load_golden("sync_codebook_2rank")
    states = [state_from(g, "pre_"), state_from(g, "pre_")]
    for step in range(2):
        xs = [T(g[f"r{r}_x{step}"]) for r in range(2)]
        # what the collective delivers: the sum over ranks of each rank's local statistic
        k = states[0]["embed"].shape[0]
        local = []
        for r in range(2):
            flat = xs[r].reshape(-1, xs[r].shape[-1])
            ind = O.neg_sq_dist(flat, states[r]["embed"]).argmax(-1)
            onehot = torch.nn.functional.one_hot(ind, k).float()
            local.append((onehot.sum(0), flat.t() @ onehot))
        total = [local[0][0] + local[1][0], local[0][1] + local[1][1]]
        for r in range(2):
            calls = iter(total)

            def hook(t, calls=calls):
                t.copy_(next(calls))
            q, ind, loss, ppl = O.vq_forward(states[r], xs[r], training=True, all_reduce=hook)
            assert np.array_equal(ind.numpy().astype(np.int16), g[f"r{r}_out{step}_ind"])
            torch.testing.assert_close(q, T(g[f"r{r}_out{step}_q"]), **RT)
            # perplexity stays rank-local in the reference
            torch.testing.assert_close(ppl, T(g[f"r{r}_out{step}_perplexity"]), **RT)
            torch.testing.assert_close(loss["loss"].detach(), T(g[f"r{r}_out{step}_loss"]), **RT)
            check_state(states[r], g, f"r{r}_post{step}_")
        for k_ in ("cluster_size", "embed_avg", "embed"):
            assert torch.equal(states[0][k_], states[1][k_]), "replicas must stay bit-identical"


def test_chunked_assignment_equals_unchunked():
    torch.manual_seed(5)
    x, e = torch.randn(5000, 64), torch.randn(300, 64)
    full = O.neg_sq_dist(x, e).argmax(-1)
    assert torch.equal(full, O.assign_chunked(x, e, max_dist_bytes=300 * 4 * 777))


def test_empty_batch_is_rejected_like_the_reference():
    """N = 0: the reference's argmax over an empty N x K matrix is fine but mean() gives NaN perplexity."""
    state = O.new_state(8, 4)
    q, ind, loss, ppl = O.vq_forward(state, torch.zeros(0, 3, 4), training=False)
    assert q.shape == (0, 3, 4) and ind.shape == (0, 3) and torch.isnan(ppl)


# ------------------------------------------------ the restatement against the LIVE reference (when it is available)

def _reference():
    try:
        import ref_loader
        return ref_loader.load()[0]
    except Exception:
        return None


@pytest.mark.parametrize("k,d,shape,kw", [(32, 128, (4, 75, 128), {}), (512, 64, (2, 300, 64), {"decay": 0.9, "commitment_weight": 0.25}),
                                          (16, 32, (3, 40, 32), {"threshold_ema_dead_code": 2}), (8, 16, (2, 100, 16), {"kmeans_init": True})])
def test_oracle_equals_live_reference(k, d, shape, kw):
    """oracle/vq_oracle.py against the UNMODIFIED timevqvae/models/vq.py (oracle/_ref, else /root/reference) on seeded inputs,
    three training steps and one eval call: indices equal, every float output and buffer bitwise equal.  Skipped only where
    neither copy of the reference exists."""
    ref_vq = _reference()
    if ref_vq is None:
        pytest.skip("reference not available (run oracle/build_ref.py where /root/reference exists)")
    torch.manual_seed(k + d)
    vq = ref_vq.VectorQuantize(d, k, **kw).train()
    cb = vq._codebook
    okw = {kk: v for kk, v in kw.items() if kk != "kmeans_init"}
    state = {n: getattr(cb, n).detach().clone() for n in ("initted", "cluster_size", "embed_avg", "embed")}
    g = torch.Generator().manual_seed(5)
    for step in range(3):
        x = torch.randn(shape, generator=g) * (1 + step)
        torch.manual_seed(50 + step)
        q, ind, loss, ppl = vq(x.clone())
        torch.manual_seed(50 + step)
        q2, ind2, loss2, ppl2 = O.vq_forward(state, x.clone(), training=True, **okw)
        assert torch.equal(ind, ind2)
        assert torch.equal(q, q2) and torch.equal(loss["loss"], loss2["loss"]) and torch.equal(ppl, ppl2)
        for n in ("cluster_size", "embed_avg", "embed", "initted"):
            assert torch.equal(getattr(cb, n).detach(), state[n]), (step, n)
    vq.eval()
    x = torch.randn(shape, generator=g)
    with torch.no_grad():
        q, ind, loss, ppl = vq(x)
    q2, ind2, loss2, ppl2 = O.vq_forward(state, x, training=False, **okw)
    assert torch.equal(ind, ind2) and torch.equal(q, q2) and torch.equal(ppl, ppl2)
