"""Experiment helper (not part of the product): time the fused train step of a given build of the library.
usage: python tools/time_variant.py <libtvq.so> [label]"""
import ctypes, sys, torch
lib = ctypes.CDLL(sys.argv[1])
label = sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
vp, i64, i, f, dbl, sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_size_t
lib.tvq_workspace_bytes.restype = sz
lib.tvq_workspace_bytes.argtypes = [i64, i, i]
lib.tvq_train_step.argtypes = [vp, vp, vp, vp, vp, i64, i, i, f, dbl, dbl, vp, vp, vp, vp, vp, vp, sz, vp]
dev = torch.device("cuda")
def graph_us(fn, reps=20):
    for j in range(3): fn(j)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g):
            for j in range(reps): fn(j)
    torch.cuda.current_stream().wait_stream(s)
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps * 1000)
    return round(min(ts), 1), round(sorted(ts)[2], 1)
k, d = 32, 128
out = {}
for n in (18432, 76800, 1 << 20):
    torch.manual_seed(0)
    xs = [torch.randn(n, d, device=dev) for _ in range(4)]
    e = torch.randn(k, d, device=dev); cs = torch.zeros(k, device=dev); avg = e.clone(); prev = e.clone()
    idx = torch.empty(n, dtype=torch.int64, device=dev); q = torch.empty(n, d, device=dev); sc = torch.empty(8, device=dev)
    wsb = lib.tvq_workspace_bytes(n, k, d); ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
    def call(j):
        st = torch.cuda.current_stream().cuda_stream
        rc = lib.tvq_train_step(xs[j % 4].data_ptr(), e.data_ptr(), cs.data_ptr(), avg.data_ptr(), prev.data_ptr(), n, k, d, 1.0, 0.8, 1e-5,
                                idx.data_ptr(), q.data_ptr(), sc.data_ptr(), None, None, ws.data_ptr(), wsb, st)
        assert rc == 0, rc
    for j in range(40): call(j)          # let the EMA settle (the codebook collapses towards the mean on random data)
    out[n] = graph_us(call)
print(label, "us per launch (best, median):", out, flush=True)
