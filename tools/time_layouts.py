"""Where the time of a quantize() call goes at BASELINE configs[1] sizes (experiment helper, not product)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1000

for (b, hw) in ((1024, 18), (1024, 75)):
    z = torch.randn(b, 128, 1, hw, device=dev)
    vq = tvq.VectorQuantize(128, 32).to(dev).eval()
    with torch.no_grad():
        t_new = graph_us(lambda: tvq.quantize(z, vq))
        def old():
            zz = z.permute(0, 2, 3, 1).reshape(b, hw, 128)
            q, i, l, p = vq(zz)
            return q.reshape(b, 1, hw, -1).permute(0, 3, 1, 2).contiguous()
        t_old = graph_us(old)
        zz = z.permute(0, 2, 3, 1).reshape(b, hw, 128).contiguous()
        t_vq = graph_us(lambda: vq(zz))
        t_tr = graph_us(lambda: tvq.glue._swap_last_two(z.view(b, 128, hw)))
    print(f"b={b} hw={hw}: quantize() eval {t_new:.1f} us (torch permute/contiguous around the module: {t_old:.1f} us; module alone {t_vq:.1f} us; one tiled transpose {t_tr:.1f} us)")
