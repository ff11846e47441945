"""compute-sanitizer target (racecheck / synccheck / memcheck): ONE small invocation of every kernel family behind the C ABI,
at sizes that finish in seconds under the tool.  usage: compute-sanitizer --tool racecheck python tools/sanitize_target.py"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tvq_b200 as tvq

dev = torch.device("cuda")
torch.manual_seed(0)
# resident-codebook tcgen05 kernel: fused train step + backward (module API), eval, channels-first call site
vq = tvq.VectorQuantize(128, 32).to(dev).train()
x = torch.randn(4, 75, 128, device=dev, requires_grad=True)
q, ind, loss, ppl = vq(x)
(q.sum() + loss["loss"].sum()).backward()
z = torch.randn(4, 128, 3, 25, device=dev, requires_grad=True)
zq, ind, loss, ppl = tvq.quantize(z, vq)
(zq.sum() + loss["loss"].sum()).backward()
vq.eval()
with torch.no_grad():
    vq(x.detach()); tvq.quantize(z.detach(), vq)
# streamed-codebook tcgen05 kernel (single CTA and CTA pairs), CUDA-core kernel
for (n, k, d) in ((300, 512, 64), (700, 1024, 128), (260, 300, 256)):
    xs, e = torch.randn(n, d, device=dev), torch.randn(k, d, device=dev)
    ws = tvq.Workspace(k, d, dev)
    tvq.vq_forward_raw(xs, e, ws, train=True)
    tvq.vq_forward_raw(xs, e, ws, train=False, flags=tvq._lib.F_NO_UMMA)
# EMA kernels, gather, dense distances, re-seed
vq2 = tvq.VectorQuantize(64, 512, threshold_ema_dead_code=2).to(dev).train()
vq2(torch.randn(2, 300, 64, device=dev))
tvq.decode_tokens(torch.randint(0, 32, (4, 75), device=dev), vq, 3, 25)
tvq.vq_neg_dist(torch.randn(100, 128, device=dev), vq._codebook.embed)
# front end, decoder-side ISTFT, MaskGIT step, Snake
xt = torch.rand(8, 4, 200, device=dev) * 2 - 1
fr = tvq.lf_hf_frontend(xt, 4)
u = torch.randn(4, 8, 3, 384, device=dev, requires_grad=True)
tvq.band_timefreq_to_time(u, 4, 4, "lf", 200).sum().backward()
tvq.maskgit_step(torch.randn(8, 75, 32, device=dev), torch.full((8, 75), 32, dtype=torch.int64, device=dev), 32, 20, 2.0)
np.random.seed(0)
act = tvq.stage1.SnakeActivation(16).to(dev)
act(torch.randn(4, 16, 3, 50, device=dev, requires_grad=True)).sum().backward()
torch.cuda.synchronize()
print("sanitize target ok")
