"""Experiment helper: the d = 256 points of the configs[2] sweep at N = 2^20 (forward kernel incl. preparation)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tvq_b200 as tvq
dev = torch.device("cuda")
HBM, TC = 6555.2e9, 1621.8e12
n = 1 << 20
for (k, d) in [(512, 256), (1024, 256), (2048, 256), (4096, 256), (8192, 256), (1024, 64), (2048, 64), (8192, 128)]:
    g = torch.Generator(device=dev).manual_seed(1)
    xs = [torch.randn(n, d, device=dev, generator=g) for _ in range(2)]
    e = torch.randn(k, d, device=dev, generator=g)
    ws = tvq.Workspace(k, d, dev)
    for i in range(3): tvq.vq_forward_raw(xs[i % 2], e, ws, train=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(6): tvq.vq_forward_raw(xs[i % 2], e, ws, train=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 6
    frac = max(n * (8 * d + 8) / HBM, 2.0 * n * k * d / TC) / (ms * 1e-3)
    sl = xs[1][:4096].contiguous()
    i1, _, _ = tvq.vq_forward_raw(sl, e, ws, train=False, write_q=False)
    i2, _, _ = tvq.vq_forward_raw(sl, e, ws, train=False, write_q=False, flags=tvq._lib.F_NO_UMMA)
    print(f"{k}x{d}: {ms:.3f} ms frac {frac:.3f} exact {bool(torch.equal(i1, i2))}", flush=True)
