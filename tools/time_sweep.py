"""Experiment helper (not part of the product): the BASELINE configs[2] sweep points (train forward incl. EMA statistics)
for a given build of the library.  usage: python tools/time_sweep.py [libtvq.so]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tvq_b200 as tvq
import tvq_b200._lib as _L
if len(sys.argv) > 1:
    _L.LIB_PATH = sys.argv[1]
dev = torch.device("cuda")
out = []
for (n, k, d) in ((1 << 22, 512, 64), (1 << 21, 1024, 128), (1 << 21, 2048, 128), (1 << 20, 4096, 128), (1 << 20, 4096, 256), (1 << 19, 16384, 256)):
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
    r = sc.view(torch.int32)[4:6].tolist()
    out.append(f"{k}x{d}: {e0.elapsed_time(e1) / 6:.3f} ms (rescored {r[0] / n:.3f}, fp64 {r[1] / n:.5f})")
print(os.path.basename(_L.LIB_PATH), " | ".join(out), flush=True)
