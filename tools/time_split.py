"""Experiment: the configs[1] VQ step (LF + HF, forward + backward, two streams, CUDA-graph replay) for different SM shares."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tvq_b200 as tvq
dev = torch.device("cuda")
B, TL, TH, K, D = 1024, 18, 75, 32, 128
gen = torch.Generator(device=dev).manual_seed(1)
sets = [tuple(torch.randn(B, t, D, device=dev, generator=gen).requires_grad_(r) for t, r in ((TL, True), (TH, True), (TL, False), (TH, False))) for _ in range(4)]
ones = torch.ones(1, device=dev)
side = torch.cuda.Stream()


def run(share_l, share_h, which="both", two_streams=True):
    torch.manual_seed(0)
    vq_l = tvq.VectorQuantize(D, K).to(dev).train(); vq_h = tvq.VectorQuantize(D, K).to(dev).train()
    vq_l._codebook.sm_share, vq_h._codebook.sm_share = share_l, share_h

    def step(xl, xh, gl, gh):
        cur = torch.cuda.current_stream()
        if which in ("both", "hf"):
            if two_streams:
                side.wait_stream(cur)
            with torch.cuda.stream(side if two_streams else cur):
                qh, ih, lh, ph = vq_h(xh)
                torch.autograd.grad([qh, lh["loss"]], [xh], [gh, ones])
        if which in ("both", "lf"):
            ql, il, ll, pl = vq_l(xl)
            torch.autograd.grad([ql, ll["loss"]], [xl], [gl, ones])
        if which in ("both", "hf") and two_streams:
            cur.wait_stream(side)
    s2 = torch.cuda.Stream()
    s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s2):
        for s in sets:
            step(*s)
    torch.cuda.current_stream().wait_stream(s2)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for j in range(8):
            step(*sets[j % 4])
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"shares LF {share_l} HF {share_h} {which:4s} two_streams={two_streams}: {e0.elapsed_time(e1) / 40 * 1e3:.1f} us per step", flush=True)


for which in ("lf", "hf"):
    for sh in (None, 0.2, 0.5, 0.8):
        run(sh, sh, which)
for sl, sh in ((None, None), (0.19, 0.81), (0.25, 0.75), (0.3, 0.7), (0.15, 0.85), (0.4, 0.6)):
    run(sl, sh)
run(None, None, "both", False)
