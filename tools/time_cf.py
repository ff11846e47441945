"""Experiment helper (not part of the product): channels-first entry points vs the row-major ones, per launch in a CUDA graph."""
import os, sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import tvq_b200 as tvq
import tvq_b200._lib as _L
if len(sys.argv) > 1:
    _L.LIB_PATH = sys.argv[1]
from tvq_b200 import functional as TF
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
dev = torch.device("cuda")
lib = tvq._lib.load()
for (b, hw) in ((1024, 18), (1024, 75), (8192, 75)):
    d, k = 128, 32
    n = b * hw
    z = torch.randn(b, d, hw, device=dev); x = torch.randn(n, d, device=dev); g = torch.randn(b, d, hw, device=dev); gr = torch.randn(n, d, device=dev)
    vq = tvq.VectorQuantize(d, k).to(dev).train(); cb = vq._codebook; ws = cb._workspace(dev)
    prev = cb.embed.detach().clone(); idx = torch.zeros(n, dtype=torch.int64, device=dev)
    q = torch.empty(n, d, device=dev); qc = torch.empty(b, d, hw, device=dev); sc = torch.empty(8, device=dev)
    one = torch.ones(1, device=dev); gz = torch.empty(b, d, hw, device=dev)
    st = lambda: torch.cuda.current_stream().cuda_stream
    E = (cb.embed.data_ptr(), cb.cluster_size.data_ptr(), cb.embed_avg.data_ptr(), prev.data_ptr())
    r = {}
    r["transpose"] = graph_us(lambda: TF.transpose12(z))
    r["train_step rowmajor"] = graph_us(lambda: lib.tvq_train_step(x.data_ptr(), *E, n, k, d, 1.0, 0.8, 1e-5, idx.data_ptr(), q.data_ptr(), sc.data_ptr(), None, None, ws.buf.data_ptr(), ws.nbytes, st()))
    r["train_step qcf"] = graph_us(lambda: lib.tvq_train_step_qcf(x.data_ptr(), *E, n, k, d, 1.0, 0.8, 1e-5, idx.data_ptr(), qc.data_ptr(), sc.data_ptr(), None, None, ws.buf.data_ptr(), ws.nbytes, None, 0, 1, hw, st()))
    r["train_step cf"] = graph_us(lambda: lib.tvq_train_step_cf(z.data_ptr(), *E, b, hw, k, d, 1.0, 0.8, 1e-5, idx.data_ptr(), qc.data_ptr(), sc.data_ptr(), None, None, ws.buf.data_ptr(), ws.nbytes, None, 0, 1, st()))
    r["backward rowmajor"] = graph_us(lambda: lib.tvq_backward(gr.data_ptr(), None, one.data_ptr(), x.data_ptr(), idx.data_ptr(), prev.data_ptr(), n, k, d, 1.0, q.data_ptr(), st()))
    r["backward cf"] = graph_us(lambda: lib.tvq_backward_cf(g.data_ptr(), None, one.data_ptr(), x.data_ptr(), idx.data_ptr(), prev.data_ptr(), b, hw, k, d, 1.0, gz.data_ptr(), st()))
    r["backward cfx"] = graph_us(lambda: lib.tvq_backward_cfx(g.data_ptr(), None, one.data_ptr(), z.data_ptr(), idx.data_ptr(), prev.data_ptr(), b, hw, k, d, 1.0, gz.data_ptr(), st()))
    print(b, hw, r, flush=True)
