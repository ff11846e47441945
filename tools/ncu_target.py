"""Small fixed workload for ncu captures: the fused forward on the product library.
usage: python tools/ncu_target.py [n] [k] [d] [train(0/1; 2 = fused train step, forward + EMA in one launch)] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tvq_b200 as tvq
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
k = int(sys.argv[2]) if len(sys.argv) > 2 else 32
d = int(sys.argv[3]) if len(sys.argv) > 3 else 128
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 1
train = mode != 0
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 4
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randn(n, d, device=dev, generator=g)
e = torch.randn(k, d, device=dev, generator=g)
ws = tvq.Workspace(k, d, dev)
if mode == 2:
    vq = tvq.VectorQuantize(d, k).to(dev).train()
    cb = vq._codebook
    ws = cb._workspace(dev)
    prev = torch.empty_like(cb.embed)
for _ in range(reps):
    if mode == 2:
        idx, q, sc, _, _ = tvq.vq_train_step_raw(x, cb, ws, 1.0, prev)
    else:
        idx, q, sc = tvq.vq_forward_raw(x, e, ws, train=train)
torch.cuda.synchronize()
print("ok", int(idx[:8].sum()), float(sc[1]), sc.view(torch.int32)[4:6].tolist())
