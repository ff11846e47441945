"""Experiment helper: configs[2] sweep points at N = 2^20 (train forward incl. EMA statistics) for the current build,
with the roofline fraction of each.  usage: python tools/time_sweep2.py [small]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tvq_b200 as tvq
dev = torch.device("cuda")
HBM, TC = 6555.2e9, 1621.8e12
pts = [(512, 64), (512, 128), (1024, 128), (2048, 128), (4096, 64), (4096, 128), (4096, 256), (16384, 64), (16384, 256)]
if len(sys.argv) > 1 and sys.argv[1] == "conv":      # the shapes whose floor is (was) the converter warps
    pts = [(512, 64), (1024, 64), (512, 128), (1024, 128), (2048, 128), (512, 256), (1024, 256), (2048, 256), (4096, 256), (16384, 256)]
if len(sys.argv) > 1 and sys.argv[1] == "small":
    pts = [(512, 64), (1024, 128), (4096, 128), (16384, 256)]
n = 1 << 20
out = []
for (k, d) in pts:
    g = torch.Generator(device=dev).manual_seed(1)
    xs = [torch.randn(n, d, device=dev, generator=g) for _ in range(2)]
    e = torch.randn(k, d, device=dev, generator=g)
    ws = tvq.Workspace(k, d, dev)
    for i in range(3):
        idx, q, sc = tvq.vq_forward_raw(xs[i % 2], e, ws, train=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(6):
        idx, q, sc = tvq.vq_forward_raw(xs[i % 2], e, ws, train=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 6
    r = sc.view(torch.int32)[4:6].tolist()
    frac = max(n * (8 * d + 8) / HBM, 2.0 * n * k * d / TC) / (ms * 1e-3)
    # exactness spot check against the CUDA-core path on a slab
    sl = xs[1][:4096].contiguous()
    i1, _, _ = tvq.vq_forward_raw(sl, e, ws, train=False, write_q=False)
    i2, _, _ = tvq.vq_forward_raw(sl, e, ws, train=False, write_q=False, flags=tvq._lib.F_NO_UMMA)
    ok = bool(torch.equal(i1, i2))
    out.append(f"{k}x{d}: {ms:.3f} ms frac {frac:.3f} (rescored {r[0] / n:.3f}, fp64 {r[1] / n:.5f}, exact {ok})")
print("\n".join(out), flush=True)
