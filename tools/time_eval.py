"""Experiment helper: eval-mode timings of the streamed path (assignment only / with q) at 2^20 latents."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tvq_b200 as tvq
dev = torch.device("cuda")
n = 1 << 20
for (k, d) in [(512, 64), (1024, 128), (4096, 128), (16384, 64)]:
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(n, d, device=dev, generator=g); e = torch.randn(k, d, device=dev, generator=g)
    ws = tvq.Workspace(k, d, dev)
    out = []
    for wq in (False, True):
        for _ in range(2): r = tvq.vq_forward_raw(x, e, ws, train=False, write_q=wq)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): r = tvq.vq_forward_raw(x, e, ws, train=False, write_q=wq)
        e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / 5)
    i_ref, _, _ = tvq.vq_forward_raw(x[:8192].contiguous(), e, ws, train=False, write_q=False, flags=tvq._lib.F_NO_UMMA)
    i_a, _, _ = tvq.vq_forward_raw(x[:8192].contiguous(), e, ws, train=False, write_q=False)
    i_b, q_b, _ = tvq.vq_forward_raw(x[:8192].contiguous(), e, ws, train=False, write_q=True)
    ok = bool(torch.equal(i_a, i_ref) and torch.equal(i_b, i_ref) and torch.equal(q_b, e[i_ref]))
    print(f"{k}x{d}: assign {out[0]:.3f} ms, assign + q {out[1]:.3f} ms, exact {ok}", flush=True)
