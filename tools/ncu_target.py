"""Small fixed workload for ncu captures: the fused forward on the product library.
usage: python tools/ncu_target.py [n] [k] [d] [train(0/1)] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tvq_b200 as tvq
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
k = int(sys.argv[2]) if len(sys.argv) > 2 else 32
d = int(sys.argv[3]) if len(sys.argv) > 3 else 128
train = bool(int(sys.argv[4])) if len(sys.argv) > 4 else True
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 4
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randn(n, d, device=dev, generator=g)
e = torch.randn(k, d, device=dev, generator=g)
ws = tvq.Workspace(k, d, dev)
for _ in range(reps):
    idx, q, sc = tvq.vq_forward_raw(x, e, ws, train=train)
torch.cuda.synchronize()
print("ok", int(idx[:8].sum()), float(sc[1]), sc.view(torch.int32)[4:6].tolist())
