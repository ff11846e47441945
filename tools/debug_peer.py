import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tvq_b200 as tvq
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
vq = tvq.VectorQuantize(128, 32, sync_codebook=True).to(dev).train()
ref = tvq.VectorQuantize(128, 32, sync_codebook=True).to(dev).train()
ref.load_state_dict(vq.state_dict())
ref._codebook._px = False
g = torch.Generator(device=dev).manual_seed(10 + rank)
for step in range(4):
    x = torch.randn(8, 75, 128, device=dev, generator=g)
    q, i, l, p = vq(x)
    q2, i2, l2, p2 = ref(x)
    torch.cuda.synchronize()
    px = vq._codebook._px
    msg = f"rank {rank} step {step}: idx_eq={torch.equal(i, i2)} q_eq={torch.equal(q, q2)}"
    for name in ("cluster_size", "embed_avg", "embed"):
        a, b = getattr(vq._codebook, name), getattr(ref._codebook, name)
        msg += f" {name}: maxdiff={float((a-b).abs().max()):.3e}"
    if px:
        msg += f" offset={px.handle.offset} ptrs={[hex(int(v)) for v in px.handle.buffer_ptrs]} buf={hex(px.buf.data_ptr())} counter={int(px.buf[:4].view(torch.int32)[0])}"
    print(msg, flush=True)
dist.barrier()
os._exit(0)
