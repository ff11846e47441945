"""Experiment helper (not part of the product): fraction of rows that leave the tensor-core level (scalars[4]) and the
launch time of the fused train step, for a fresh codebook and after many EMA steps on random data."""
import os, sys, torch
sys.path.insert(0, "/root/repo")
import tvq_b200 as tvq
dev = torch.device("cuda")
def graph_us(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g):
            for _ in range(reps): fn()
    torch.cuda.current_stream().wait_stream(s)
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps * 1000)
    return round(best, 1)
n, d, k = 76800, 128, 32
xs = [torch.randn(n, d, device=dev) for _ in range(4)]
vq = tvq.VectorQuantize(d, k).to(dev).train(); cb = vq._codebook; ws = cb._workspace(dev)
prev = torch.empty_like(cb.embed)
def frac():
    e = cb.embed.detach().clone()
    idx, q, sc = tvq.vq_forward_raw(xs[0], e, tvq.Workspace(k, d, dev), train=True)
    torch.cuda.synchronize()
    r = sc.view(torch.int32)[4:6].tolist()
    return r[0] / n, r[1] / n, float(e.norm(dim=1).max())
state = {"i": 0}
def step():
    tvq.vq_train_step_raw(xs[state["i"] % 4], cb, ws, 1.0, prev); state["i"] += 1
def eval_us():
    e = cb.embed.detach().clone(); w2 = tvq.Workspace(k, d, dev)
    return graph_us(lambda: tvq.vq_forward_raw(xs[0], e, w2, train=True))
print("fresh codebook: rescored, fp64, max|e| =", frac(), " forward(prep+fwd) us:", eval_us())
for steps in (5, 50, 300):
    while state["i"] < steps: step()
    print(f"after {steps} EMA steps: rescored, fp64, max|e| =", frac(), " forward(prep+fwd) us:", eval_us())
