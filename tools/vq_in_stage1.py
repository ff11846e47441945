"""How the VQ kernels behave on REAL stage-1 latents (untrained and after some optimisation steps): time per launch,
rows that left the tensor-core nomination level (scalars[4]) / needed the full scan (scalars[5]), norms of z and the codes."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tvq_b200 as tvq

dev = torch.device("cuda", 0)
torch.backends.cudnn.allow_tf32 = True
torch.manual_seed(0); np.random.seed(0)
B = 1024
model = tvq.Stage1(200, 4, tvq.stage1.default_config()).to(dev)
tr = tvq.Stage1Trainer(model, (B, 4, 200), use_graph=False)
gen = torch.Generator(device=dev).manual_seed(1)


def smooth_batch():
    # trajectory-like: smooth random curves scaled to [-1, 1]
    t = torch.linspace(0, 1, 200, device=dev)
    f = torch.rand(B, 4, 3, 1, device=dev, generator=gen) * 6
    ph = torch.rand(B, 4, 3, 1, device=dev, generator=gen) * 6.28
    a = torch.rand(B, 4, 3, 1, device=dev, generator=gen)
    x = (a * torch.sin(f * t * 6.28 + ph)).sum(2)
    return x / x.abs().amax(dim=(1, 2), keepdim=True).clamp_min(1e-6)


def probe(tag):
    x = smooth_batch()
    with torch.no_grad():
        fr = tvq.lf_hf_frontend(x, 4, want=("enc_in_l", "enc_in_h"))
        for name, enc, vq in (("LF", model.encoder_l, model.vq_model_l), ("HF", model.encoder_h, model.vq_model_h)):
            z = enc(fr["enc_in_l" if name == "LF" else "enc_in_h"])
            b, c, h, w = z.shape
            flat = z.permute(0, 2, 3, 1).reshape(-1, c).contiguous()
            cb = vq._codebook
            ws = tvq.Workspace(32, 128, dev)
            e = cb.embed.clone()
            for _ in range(3):
                idx, q, sc = tvq.vq_forward_raw(flat, e, ws, train=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                idx, q, sc = tvq.vq_forward_raw(flat, e, ws, train=True)
            e1.record(); torch.cuda.synchronize()
            d = sc.view(torch.int32)
            mu = e.mean(0)
            print(f"{tag} {name}: n={flat.shape[0]} {e0.elapsed_time(e1) / 10 * 1e3:.1f} us/launch, fp64-rescored {int(d[4])} full-scan {int(d[5])}, "
                  f"|z| mean {float(flat.norm(dim=1).mean()):.2f}, |z - mean_z| {float((flat - flat.mean(0)).norm(dim=1).mean()):.2f}, "
                  f"|e| {float(e.norm(dim=1).mean()):.2f}, |e - mean_e| {float((e - mu).norm(dim=1).mean()):.2f}, "
                  f"|mean_e| {float(mu.norm()):.2f}, codes used {int(torch.bincount(idx, minlength=32).gt(0).sum())}, ppl {float(sc[1]):.2f}", flush=True)


probe("step 0")
for it in range(1, 301):
    tr.step(smooth_batch())
    if it in (20, 100, 300):
        probe(f"step {it}")
