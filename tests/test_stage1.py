"""Stage-1 harness (t-vq-vae-trajgen_b200/stage1.py) against the UNMODIFIED reference Stage1 (tests/golden/stage1_cfg0.npz,
oracle/gen_golden_stage1.py): BASELINE configs[0] — Stage1(200, 4, configs/config.yaml), batch 32 x 4 x 200.

CPU: the harness reproduces the reference's initialisation from the same two seeds (all 454 state tensors, same names, same
order) and its learning-rate schedule.  GPU: one training step (forward, backward, AdamW at the schedule's first learning
rate) — losses, perplexities, gradients, post-step parameters, EMA buffers.  Tolerances: the conv stacks run as cuDNN fp32
here and as MKL/oneDNN fp32 in the fixture, so activations agree to ~1e-6 relative and a latent on a decision boundary may
take the neighbouring code; scalars are compared at 1e-4, gradients at 2e-3 of their norm.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden


def build(tvq, device="cpu", dropout_off=True):
    torch.manual_seed(0)
    np.random.seed(0)
    model = tvq.stage1.Stage1(200, 4, tvq.stage1.default_config())
    if dropout_off:
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
    return model.to(device).train()


@pytest.fixture(scope="module")
def tvq():
    import tvq_b200
    return tvq_b200


def test_harness_reproduces_reference_initialisation(tvq):
    g = load_golden("stage1_cfg0")
    sd = build(tvq).state_dict()
    names = [str(n) for n in g["state_names"]]
    assert list(sd.keys()) == names, "state-dict keys (or their order) differ from the reference's Stage1"
    for n, s, a, numel in zip(names, g["state_sum"], g["state_abs"], g["state_numel"]):
        t = sd[n].double()
        assert t.numel() == int(numel), n
        assert float(t.sum()) == pytest.approx(float(s), rel=1e-12, abs=1e-12), n
        assert float(t.abs().sum()) == pytest.approx(float(a), rel=1e-12, abs=1e-12), n
    assert sum(p.numel() for p in build(tvq).parameters()) == 1_112_928


def test_learning_rate_schedule(tvq):
    g = load_golden("stage1_lr")
    for step, lr in zip(g["steps"], g["lr"]):
        f = tvq.stage1.warmup_cosine_factor(int(step), int(g["max_steps"]), float(g["warmup_rate"]), float(g["base_lr"]))
        assert f * float(g["base_lr"]) == pytest.approx(float(lr), rel=1e-9, abs=1e-15), int(step)


def test_token_grid_of_the_shipped_config(tvq):
    from tvq_b200.stage1 import compute_downsample_rate
    assert compute_downsample_rate(200, 4, 8) == 25 and compute_downsample_rate(200, 4, 32) == 6    # -> 18 LF / 75 HF tokens


@pytest.mark.gpu
def test_one_training_step_matches_reference(tvq):
    g = load_golden("stage1_cfg0")
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        model = build(tvq, "cuda")
        x = torch.from_numpy(g["x"]).cuda()
        trainer = tvq.stage1.Stage1Trainer(model, x.shape, use_graph=False)
        out = trainer.step(x)
        torch.cuda.synchronize()
        assert float(trainer.lr) == pytest.approx(float(g["lr_step1"]), rel=1e-6)
        for key, ref in (("loss", "loss"), ("recons_loss.LF.time", "recons_lf"), ("recons_loss.HF.time", "recons_hf"),
                         ("perplexity.LF", "ppl_lf"), ("perplexity.HF", "ppl_hf")):
            assert float(out[key].reshape(-1)[0]) == pytest.approx(float(np.asarray(g[ref]).reshape(-1)[0]), rel=1e-4), key
        names = [str(n) for n in g["param_names"]]
        params = dict(model.named_parameters())
        assert list(params.keys()) == names
        norms = np.array([float(params[n].grad.double().norm()) for n in names])
        ref = g["grad_norm"]
        big = ref > 1e-3 * ref.max()
        np.testing.assert_allclose(norms[big], ref[big], rtol=2e-3)
        for key in list(g.keys()):
            if key.startswith("grad::"):
                r = torch.from_numpy(g[key])
                torch.testing.assert_close(params[key[6:]].grad.cpu(), r, rtol=0, atol=2e-3 * float(r.abs().max()) + 1e-9, msg=lambda m: f"{key}: {m}")
            if key.startswith("post::"):
                r = torch.from_numpy(g[key])
                # first AdamW step: |update| = lr wherever |g| >> eps, so compare to a fraction of the step
                torch.testing.assert_close(params[key[6:]].detach().cpu(), r, rtol=0, atol=0.1 * float(g["lr_step1"]) + 1e-7,
                                           msg=lambda m: f"{key}: {m}")
        sd = model.state_dict()
        for key in list(g.keys()):
            if key.startswith("poststate::"):
                r = torch.from_numpy(g[key])
                torch.testing.assert_close(sd[key[11:]].cpu(), r, rtol=1e-3, atol=1e-3 * float(r.abs().max()), msg=lambda m: f"{key}: {m}")
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.gpu
def test_graph_replay_equals_eager_steps(tvq):
    """Three optimisation steps replayed from the captured CUDA graph == the same three steps run eagerly (dropout off)."""
    xs = [torch.rand(64, 4, 200, device="cuda", generator=torch.Generator(device="cuda").manual_seed(s)) * 2 - 1 for s in range(3)]
    res = []
    for use_graph in (False, True):
        model = build(tvq, "cuda")
        tr = tvq.stage1.Stage1Trainer(model, xs[0].shape, use_graph=use_graph)
        if use_graph:
            sd0 = {k: v.clone() for k, v in model.state_dict().items()}
            tr.warmup_and_capture(2)
            model.load_state_dict(sd0)                      # back to the initial weights; optimizer moments restart below
            for st in tr.opt.state.values():
                for v in st.values():
                    if torch.is_tensor(v):
                        v.zero_()
            tr.step_count = 0
        losses = [float(tr.step(x)["loss"].reshape(-1)[0]) for x in xs]
        res.append((losses, {k: v.clone() for k, v in model.state_dict().items()}))
    (l0, s0), (l1, s1) = res
    np.testing.assert_allclose(l1, l0, rtol=1e-4)
    # (the codebooks are not compared element-wise: on an untrained encoder the latents sit on top of one another, and the
    # 1e-6 differences between two cuDNN algorithm choices move a few of them across a decision boundary)
    for k in ("decoder_l.linear.weight", "encoder_h.encoder.0.block.0.weight"):
        torch.testing.assert_close(s1[k], s0[k], rtol=1e-3, atol=1e-5)
    assert float(s1["vq_model_h._codebook.cluster_size"].sum()) == pytest.approx(float(s0["vq_model_h._codebook.cluster_size"].sum()), rel=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,cl", [((3, 8, 3, 201), False), ((3, 8, 3, 201), True), ((5, 128, 3, 25), True), ((2, 4, 1, 7), False),
                                      ((64, 16, 3, 100), True)])
def test_fused_snake_matches_torch_expression(tvq, shape, cl):
    """SnakeActivation on the fused kernels vs the reference's expression x + (1 / a) * sin(a x)^2 under torch autograd
    (utils/train_utils.py:447), both memory formats: forward 1e-6, gradients 1e-5 (the gradient of `a` is a sum over
    up to 1e5 elements accumulated with atomics)."""
    from tvq_b200.stage1 import SnakeActivation
    g = torch.Generator(device="cuda").manual_seed(5)
    np.random.seed(1)
    act = SnakeActivation(shape[1]).cuda()
    x = torch.randn(shape, device="cuda", generator=g) * 3
    if cl:
        x = x.contiguous(memory_format=torch.channels_last)
    gy = torch.randn(shape, device="cuda", generator=g)
    x1 = x.clone().requires_grad_(True)
    y1 = act(x1)
    y1.backward(gy)
    ga1 = act.a.grad.clone(); act.a.grad = None
    x2 = x.clone().requires_grad_(True)
    y2 = x2 + (1 / act.a) * torch.sin(act.a * x2) ** 2
    y2.backward(gy)
    torch.testing.assert_close(y1, y2, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(x1.grad, x2.grad, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(ga1, act.a.grad, rtol=1e-4, atol=1e-4 * float(act.a.grad.abs().max()))
