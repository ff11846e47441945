"""Experiment helper: the per-call preparation kernel and the EMA update kernel of the streamed path, alone."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tvq_b200 as tvq
from tvq_b200 import functional as TF
dev = torch.device("cuda")
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1000
for (k, d) in [(512, 64), (4096, 128), (16384, 64), (16384, 256)]:
    vq = tvq.VectorQuantize(d, k).to(dev).train(); cb = vq._codebook; ws = cb._workspace(dev)
    x = torch.randn(256, d, device=dev)
    fwd = t(lambda: tvq.vq_forward_raw(x, cb.embed, ws, train=True))
    ema = t(lambda: TF.vq_ema_update(ws.stats, cb.cluster_size, cb.embed_avg, cb.embed, None, 0.8, 1e-5, ws))
    print(k, d, f"prep + 256-latent forward {fwd:.1f} us, ema {ema:.1f} us", flush=True)
