"""Experiment helper (not part of the product): per-launch time of the fused forward at small n, back to back in a CUDA graph."""
import os, sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
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
lib = tvq._lib.load()
d, k = 128, 32
empty = torch.empty(1, device=dev)
print("empty torch kernel:", graph_us(lambda: empty.add_(1.0)))
for n in (64, 148 * 64, 18432, 76800, 4 * 76800):
    x = torch.randn(n, d, device=dev); gr = torch.randn(n, d, device=dev)
    vq = tvq.VectorQuantize(d, k).to(dev).train(); cb = vq._codebook; ws = cb._workspace(dev)
    prev = torch.empty_like(cb.embed); idx = torch.empty(n, dtype=torch.int64, device=dev)
    q = torch.empty(n, d, device=dev); sc = torch.empty(8, device=dev); stats = torch.zeros(64 + k * d, device=dev)
    one = torch.ones(1, device=dev)
    st = lambda: torch.cuda.current_stream().cuda_stream
    r = {}
    r["train_step"] = graph_us(lambda: lib.tvq_train_step(x.data_ptr(), cb.embed.data_ptr(), cb.cluster_size.data_ptr(), cb.embed_avg.data_ptr(), prev.data_ptr(), n, k, d, 1.0, 0.8, 1e-5, idx.data_ptr(), q.data_ptr(), sc.data_ptr(), None, None, ws.buf.data_ptr(), ws.nbytes, st()))
    r["eval_fwd(prep+fwd)"] = graph_us(lambda: lib.tvq_forward(x.data_ptr(), cb.embed.data_ptr(), n, k, d, 2, 1.0, idx.data_ptr(), q.data_ptr(), stats.data_ptr(), sc.data_ptr(), ws.buf.data_ptr(), ws.nbytes, st()))
    r["backward"] = graph_us(lambda: lib.tvq_backward(gr.data_ptr(), None, one.data_ptr(), x.data_ptr(), idx.data_ptr(), prev.data_ptr(), n, k, d, 1.0, q.data_ptr(), st()))
    def both():
        lib.tvq_train_step(x.data_ptr(), cb.embed.data_ptr(), cb.cluster_size.data_ptr(), cb.embed_avg.data_ptr(), prev.data_ptr(), n, k, d, 1.0, 0.8, 1e-5, idx.data_ptr(), q.data_ptr(), sc.data_ptr(), None, None, ws.buf.data_ptr(), ws.nbytes, st())
        lib.tvq_backward(gr.data_ptr(), None, one.data_ptr(), x.data_ptr(), idx.data_ptr(), prev.data_ptr(), n, k, d, 1.0, q.data_ptr(), st())
    r["train_step+backward"] = graph_us(both)
    print(n, r, flush=True)
