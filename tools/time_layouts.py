"""Where the time of a quantize() call goes at BASELINE configs[1] sizes (experiment helper, not product):
the reference-shaped glue (torch permute / contiguous around the module) against tvq_b200.quantize()."""
import os, sys, json, torch
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
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps * 1000)
    return best

def reference_glue(z, vq):
    b, c, h, w = z.shape
    zz = z.permute(0, 2, 3, 1).reshape(b, h * w, c)               # rearrange 'b c h w -> b (h w) c'
    q, i, l, p = vq(zz)
    return q.reshape(b, h, w, -1).permute(0, 3, 1, 2), i, l, p    # rearrange back (a view, as einops gives)

out = {}
for (b, hw) in ((1024, 18), (1024, 75)):
    z = torch.randn(b, 128, 1, hw, device=dev)
    g = torch.randn(b, 128, 1, hw, device=dev)
    ones = torch.ones(1, device=dev)
    vq = tvq.VectorQuantize(128, 32).to(dev)
    res = {}
    for mode in ("eval", "train"):
        vq.train(mode == "train")
        def run(glue):
            zz = z.detach().requires_grad_(mode == "train")
            with torch.set_grad_enabled(mode == "train"):
                zq, i, l, p = glue(zz, vq)
                if mode == "train":
                    torch.autograd.grad([zq, l["loss"]], [zz], [g, ones])
        res[mode] = {"tvq_quantize_us": graph_us(lambda: run(tvq.quantize)), "reference_glue_us": graph_us(lambda: run(reference_glue))}
    out[f"b{b}_hw{hw}"] = res
    print(f"b={b} hw={hw}:", json.dumps(res))
