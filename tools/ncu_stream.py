"""ncu target: a few launches of the streamed-codebook forward at one shape.  usage: ncu_stream.py n k d"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tvq_b200 as tvq
n, k, d = (int(a) for a in sys.argv[1:4])
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randn(n, d, device=dev, generator=g)
e = torch.randn(k, d, device=dev, generator=g)
ws = tvq.Workspace(k, d, dev)
for i in range(3):
    idx, q, sc = tvq.vq_forward_raw(x, e, ws, train=True)
torch.cuda.synchronize()
print("ok", int(idx[0]))
