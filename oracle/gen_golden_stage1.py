"""Generate tests/golden/stage1_*.npz and istft_dec_*.npz by running the UNMODIFIED reference in this container
(/root/reference/timevqvae/trainers/stage1.py, models/vq_vae.py, utils/train_utils.py).  TEST INFRASTRUCTURE ONLY.

    python oracle/gen_golden_stage1.py

stage1_cfg0     BASELINE configs[0] / SURVEY section 8(d) config 1: Stage1(200, 4, configs/config.yaml) built under
                torch.manual_seed(0); np.random.seed(0), batch 32 x 4 x 200 of U(-1, 1), TRAIN mode forward + backward + one
                AdamW step at the schedule's first learning rate.  Dropout is the only RNG consumer of the forward and CPU /
                CUDA generators differ, so every nn.Dropout of the (otherwise unmodified) model is set to p = 0 for the
                fixture.  Stored: checksums of all 454 initial state tensors (the harness must reproduce the reference's
                initialisation from the same two seeds), losses, perplexities, token histograms, per-parameter gradient
                norms, a few full gradients and post-step parameters.
stage1_lr       linear_warmup_cosine_annealingLR (utils/train_utils.py:451-472) sampled over a 50 000-step run.
istft_dec_*     the decoder tail F.interpolate(timefreq_to_time(pad_func(u))) for the decoders' real output widths
                (384 / 400 frames -> 200 samples), forward and autograd backward.
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
import yaml

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_golden as G

OUT = G.OUT


def stage1_case():
    _, ref_stage1, tu = G.load_reference()
    cfg = yaml.safe_load(open(os.path.join(G.REF, "configs", "config.yaml")))
    torch.manual_seed(0)
    np.random.seed(0)
    model = ref_stage1.Stage1(200, 4, cfg)
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
    model.train()
    sd = model.state_dict()
    names = list(sd.keys())
    rec = {"state_names": np.array(names), "state_sum": np.array([float(sd[k].double().sum()) for k in names]),
           "state_abs": np.array([float(sd[k].double().abs().sum()) for k in names]),
           "state_numel": np.array([sd[k].numel() for k in names])}
    x = torch.rand(32, 4, 200, generator=torch.Generator().manual_seed(0)) * 2 - 1
    y = torch.zeros(32, 1, dtype=torch.int64)
    rec["x"] = x.numpy()
    opt = torch.optim.AdamW(model.parameters(), lr=cfg["exp_params"]["lr"])
    sch = tu.linear_warmup_cosine_annealingLR(opt, cfg["trainer_params"]["max_steps"]["stage1"],
                                              cfg["exp_params"]["linear_warmup_rate"])
    # the pieces of Stage1.forward that the fixture also wants to see (same calls, same order: trainers/stage1.py:115-123)
    with torch.no_grad():
        z_l = model.encoder_l(x)
        z_h = model.encoder_h(x)
    # BatchNorm running statistics were touched by the probe above: rebuild the model so the recorded step is the first
    torch.manual_seed(0)
    np.random.seed(0)
    model = ref_stage1.Stage1(200, 4, cfg)
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
    model.train()
    opt = torch.optim.AdamW(model.parameters(), lr=cfg["exp_params"]["lr"])
    sch = tu.linear_warmup_cosine_annealingLR(opt, cfg["trainer_params"]["max_steps"]["stage1"],
                                              cfg["exp_params"]["linear_warmup_rate"])
    recons, vql, ppl = model((x, y), batch_idx=1)
    loss = (recons["LF.time"] + recons["HF.time"]) + vql["LF"]["loss"] + vql["HF"]["loss"]
    sch.step()                                                     # training_step: scheduler first (stage1.py:178-179)
    rec["lr_step1"] = np.float64(opt.param_groups[0]["lr"])
    opt.zero_grad()
    loss.backward()
    pnames = [n for n, _ in model.named_parameters()]
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
    rec.update({"z_l0": z_l[0].numpy(), "z_h0": z_h[0].numpy(),
                "loss": loss.detach().numpy(), "recons_lf": recons["LF.time"].detach().numpy(),
                "recons_hf": recons["HF.time"].detach().numpy(), "vq_loss_lf": vql["LF"]["loss"].detach().numpy(),
                "vq_loss_hf": vql["HF"]["loss"].detach().numpy(), "ppl_lf": ppl["LF"].numpy(), "ppl_hf": ppl["HF"].numpy(),
                "param_names": np.array(pnames), "grad_norm": np.array([float(grads[n].double().norm()) for n in pnames])})
    keep = ["encoder_l.encoder.0.block.0.weight", "encoder_h.encoder.5.convs.1.weight", "decoder_l.linear.weight",
            "decoder_h.decoder.0.convs.0.a", "decoder_h.decoder.8.bias"]
    for n in keep:
        assert n in grads, (n, pnames[:40])
        rec["grad::" + n] = grads[n].numpy()
    opt.step()
    post = dict(model.named_parameters())
    for n in keep:
        rec["post::" + n] = post[n].detach().numpy()
    sd2 = model.state_dict()
    for k in ("vq_model_l._codebook.embed", "vq_model_h._codebook.cluster_size", "encoder_l.encoder.0.block.1.running_mean",
              "decoder_h.decoder.0.convs.2.running_var"):
        rec["poststate::" + k] = sd2[k].numpy()
    np.savez_compressed(os.path.join(OUT, "stage1_cfg0.npz"), **rec)
    print("stage1_cfg0", len(names), "state tensors, loss", float(loss), "ppl", float(ppl["LF"]), float(ppl["HF"]),
          os.path.getsize(os.path.join(OUT, "stage1_cfg0.npz")) // 1024, "KiB")


def lr_case():
    _, _, tu = G.load_reference()
    p = nn.Parameter(torch.zeros(1))
    opt = torch.optim.AdamW([p], lr=1e-3)
    sch = tu.linear_warmup_cosine_annealingLR(opt, 50000, 0.1)
    steps = sorted(set([1, 2, 3, 10, 100, 2500, 4999, 5000, 5001, 5002, 10000, 27500, 40000, 49998, 49999, 50000]))
    lrs = {}
    import warnings
    warnings.filterwarnings("ignore")
    for s in range(1, 50001):
        opt.step()
        sch.step()
        if s in steps:
            lrs[s] = opt.param_groups[0]["lr"]
    np.savez_compressed(os.path.join(OUT, "stage1_lr.npz"), steps=np.array(steps), lr=np.array([lrs[s] for s in steps]),
                        max_steps=np.int64(50000), warmup_rate=np.float64(0.1), base_lr=np.float64(1e-3))
    print("stage1_lr", lrs[1], lrs[5000], lrs[50000])


def decoder_istft_cases():
    _, _, tu = G.load_reference()
    for name, (b, c, t, l, n_fft) in {"istft_dec_lf": (3, 4, 384, 200, 4), "istft_dec_hf": (3, 4, 400, 200, 4),
                                      "istft_dec_up": (2, 2, 41, 333, 8)}.items():
        g = torch.Generator().manual_seed(31 + t)
        u = torch.randn(b, 2 * c, n_fft // 2 + 1, t, generator=g)
        gy = torch.randn(b, c, l, generator=g)
        rec = {"u": u.numpy(), "g_y": gy.numpy(), "n_fft": np.int64(n_fft)}
        for band, pad in (("all", lambda v: v), ("lf", tu.zero_pad_high_freq), ("hf", tu.zero_pad_low_freq)):
            uu = u.clone().requires_grad_(True)
            y = F.interpolate(tu.timefreq_to_time(pad(uu), n_fft, c), l, mode="linear")
            (y * gy).sum().backward()
            rec["y_" + band] = y.detach().numpy()
            rec["g_u_" + band] = uu.grad.numpy()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        print(name, tuple(u.shape), "->", l)


if __name__ == "__main__":
    decoder_istft_cases()
    lr_case()
    stage1_case()
