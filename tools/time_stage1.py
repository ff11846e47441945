"""Quick timing of the stage-1 harness step (B trajectories per GPU), eager vs CUDA-graph replay; single GPU or torchrun."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tvq_b200 as tvq

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
torch.backends.cudnn.allow_tf32 = True
torch.set_float32_matmul_precision("high")          # scripts/train.py:19
torch.backends.cudnn.benchmark = True


def run(use_graph, channels_last, two_streams=True):
    tvq.Stage1.two_streams = two_streams
    torch.manual_seed(0); np.random.seed(0)
    cfg = tvq.stage1.default_config()
    if world > 1:
        cfg["VQ-VAE"]["sync_codebook"] = True
    model = tvq.Stage1(200, 4, cfg).to(dev)
    if channels_last:
        model = model.to(memory_format=torch.channels_last)
    tr = tvq.Stage1Trainer(model, (B, 4, 200), use_graph=use_graph)
    tr.warmup_and_capture(3)
    xs = [torch.rand(B, 4, 200, device=dev) * 2 - 1 for _ in range(4)]
    for i in range(5):
        tr.step(xs[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 30
    e0.record()
    for i in range(n):
        out = tr.step(xs[i % 4])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    if rank == 0:
        print(f"B={B} world={world} graph={use_graph} channels_last={channels_last} two_streams={two_streams}: {ms:.3f} ms/step = {B * world / ms * 1e3:,.0f} traj/s, "
              f"loss {float(out['loss'].reshape(-1)[0]):.4f}", flush=True)


for g, cl, ts in ((True, True, False), (True, True, True), (False, True, True)):
    try:
        run(g, cl, ts)
    except Exception as e:
        print("FAILED", g, cl, repr(e)[:500], flush=True)
if world > 1:
    torch.cuda.synchronize(); dist.barrier(); dist.destroy_process_group()
