"""Multi-GPU worker of tests/test_dp_peer_gpu.py (and of `gpurun --gpus N` runs kept under profiles/): one process per
GPU under torchrun, NCCL group for the rendezvous.  Every check goes through the SHIPPED data-parallel kernels
(tvq_train_step_dp and the channels-first tvq_train_step_qcf with peers: the statistics exchange over NVLink peer memory
inside the forward kernel's last CTA), never through a host-side sum.

  1. world == 2: tests/golden/sync_codebook_2rank.npz — the unmodified reference run on two gloo ranks with
     sync_codebook=True (oracle/gen_golden.py) — indices bit-exact, q / loss / perplexity / buffers within 1e-5.
  2. any world: BASELINE configs[3] shapes (b = 64 trajectories per rank, HF 75 and LF 18 tokens, K = 32, D = 128), three
     steps through VectorQuantize.forward AND quantize() (channels-first, requires_grad input, backward), against a
     single-process full-batch run of the non-data-parallel kernels on the concatenated batch (all-reduced statistics ==
     full-batch statistics up to summation order): indices bit-exact, gradient / buffers within 1e-5.
  3. replicas bit-identical after every step (all_gather of embed / embed_avg / cluster_size), sum(counts) == world * n.
  4. dead-code re-seeding (threshold_ema_dead_code = 2) and k-means init under data parallelism: replicas stay identical.
  5. the streamed-codebook shapes (k = 512): tvq_forward + tvq_ema_update_dp, same checks as 2.
Prints DP_OK <rank> on success."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = sys.argv[1] if len(sys.argv) > 1 else os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tvq_b200 as tvq  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def T(a):
    return torch.from_numpy(np.asarray(a))


def close(actual, expected, what, rtol=1e-5):
    expected = expected.detach().float().cpu()
    atol = max(rtol * float(expected.abs().max()), 1e-12)
    torch.testing.assert_close(actual.detach().float().cpu(), expected, rtol=rtol, atol=atol, msg=lambda m: f"[rank {rank}] {what}: {m}")


def replicas_identical(cb, what):
    for name in ("cluster_size", "embed_avg", "embed"):
        a = getattr(cb, name).detach().contiguous()
        gathered = [torch.empty_like(a) for _ in range(world)]
        dist.all_gather(gathered, a)
        assert all(torch.equal(gathered[0], t) for t in gathered), f"[rank {rank}] {what}: replicas diverged in {name}"


def set_state(vq, g, prefix):
    cb = vq._codebook
    with torch.no_grad():
        for name in ("initted", "cluster_size", "embed_avg", "embed"):
            getattr(cb, name).copy_(T(g[prefix + name]))
    cb._initted_host = None


# ---- 1. the reference's own 2-rank run ---------------------------------------------------------------------------------
if world == 2:
    g = np.load(os.path.join(ROOT, "tests", "golden", "sync_codebook_2rank.npz"))
    vq = tvq.VectorQuantize(32, 16, sync_codebook=True)
    set_state(vq, g, "pre_")
    vq = vq.to(dev).train()
    for step in range(2):
        x = T(g[f"r{rank}_x{step}"]).to(dev).requires_grad_(True)
        q, ind, loss, ppl = vq(x)
        assert vq._codebook._px, "the fused peer-exchange kernel was not used"
        assert np.array_equal(ind.cpu().numpy().astype(np.int16), g[f"r{rank}_out{step}_ind"]), f"step {step}: indices"
        if step == 0:
            assert torch.equal(q.detach().cpu(), T(g[f"r{rank}_out{step}_q"])), "step 0: q must be bit-exact"
        close(q, T(g[f"r{rank}_out{step}_q"]), f"step {step} q")
        close(loss["loss"], T(g[f"r{rank}_out{step}_loss"]), f"step {step} loss")
        close(ppl, T(g[f"r{rank}_out{step}_perplexity"]), f"step {step} perplexity")
        for name in ("cluster_size", "embed_avg", "embed"):
            close(getattr(vq._codebook, name), T(g[f"r{rank}_post{step}_{name}"]), f"step {step} {name}")
        replicas_identical(vq._codebook, f"golden step {step}")
    vq._codebook.check_peer_errors()
    if rank == 0:
        print("golden sync_codebook_2rank reproduced through tvq_train_step_dp", flush=True)


# ---- 2./3. configs[3] shapes against a single-process full-batch run ---------------------------------------------------
def full_batch_case(k, d, hw, b_per_rank, channels_first, steps=3, defer=False):
    torch.manual_seed(7)
    vq = tvq.VectorQuantize(d, k, sync_codebook=True, defer_exchange=defer).to(dev).train()       # data-parallel replica
    solo = tvq.VectorQuantize(d, k, sync_codebook=False).to(dev).train()    # full batch on this GPU alone
    solo.load_state_dict(vq.state_dict())
    gen = torch.Generator().manual_seed(1234 + k + hw)
    for step in range(steps):
        if channels_first:
            zfull = torch.randn(world * b_per_rank, d, 1, hw, generator=gen) * (1.0 + 0.3 * step)
        else:
            zfull = torch.randn(world * b_per_rank, hw, d, generator=gen) * (1.0 + 0.3 * step)
        gfull = torch.randn(zfull.shape, generator=gen)
        sl = slice(rank * b_per_rank, (rank + 1) * b_per_rank)
        z = zfull[sl].to(dev).requires_grad_(True)
        zs = zfull.to(dev).requires_grad_(True)
        if channels_first:
            q, ind, loss, ppl = tvq.quantize(z, vq)
            qs, inds, losss, ppls = tvq.quantize(zs, solo)
        else:
            q, ind, loss, ppl = vq(z)
            qs, inds, losss, ppls = solo(zs)
        assert vq._codebook._px or k > 32, "the fused peer-exchange kernel was not used"
        assert (vq._codebook.__dict__["_pending"] is not None) == (defer and k <= 32), "deferred finalize not scheduled as expected"
        # backward: DDP averages the ranks' gradients, each rank's loss is its local mean -> compare per shard with the
        # full-batch run whose commit loss is the global mean: g_local = g_q + (1/n_local) * ..., g_full = g_q + (1/n_full) * ...
        (q * gfull[sl].to(dev)).sum().backward()
        (qs * gfull.to(dev)).sum().backward()
        assert torch.equal(ind.reshape(-1), inds.reshape(world, -1)[rank]), f"step {step}: shard indices differ from the full-batch run"
        assert torch.equal(q.detach(), qs.detach()[sl]), f"step {step}: q differs"
        assert torch.equal(z.grad, zs.grad[sl]), f"step {step}: straight-through gradient differs"
        cb, cs = vq._codebook, solo._codebook
        for name in ("cluster_size", "embed_avg", "embed"):
            close(getattr(cb, name), getattr(cs, name), f"k={k} hw={hw} cf={channels_first} step {step} {name}")
        replicas_identical(cb, f"k={k} hw={hw} cf={channels_first} step {step}")
        # replicas continue from the solo state so that exact index equality can be asserted at every step
        solo.load_state_dict(vq.state_dict())
    vq._codebook.check_peer_errors()


for (k, d, hw, cf) in ((32, 128, 75, False), (32, 128, 18, False), (32, 128, 75, True), (32, 128, 18, True), (16, 64, 40, False),
                       (512, 64, 64, False)):
    full_batch_case(k, d, hw, 64, cf)
# the same with the exchange deferred to tvq_ema_finalize_dp on a side stream (tvq_hint_defer_exchange)
for mode in (1, 2):
    for (k, d, hw, cf) in ((32, 128, 75, False), (32, 128, 18, True), (16, 64, 40, False)):
        full_batch_case(k, d, hw, 64, cf, defer=mode)
if rank == 0:
    print("configs[3] shapes: fused data-parallel step == full-batch single-GPU step (indices/q/grad exact, buffers 1e-5)", flush=True)

# commit-loss gradient through the data-parallel step (local mean, as the reference under DDP)
torch.manual_seed(3)
vq = tvq.VectorQuantize(128, 32, sync_codebook=True, commitment_weight=0.5).to(dev).train()
x = torch.randn(8, 75, 128, device=dev, generator=torch.Generator(device=dev).manual_seed(50 + rank)).requires_grad_(True)
pre = vq._codebook.embed.clone()
q, ind, loss, ppl = vq(x)
loss["loss"].sum().backward()
qst = x.detach() + (pre[ind] - x.detach())
close(x.grad, 0.5 * 2.0 / x.numel() * (x.detach() - qst), "commit-loss gradient")

# ---- 4. RNG-consuming branches stay replica-consistent -----------------------------------------------------------------
torch.manual_seed(11 + rank)                                # DIFFERENT generators per rank on purpose
vq = tvq.VectorQuantize(64, 32, sync_codebook=True, threshold_ema_dead_code=2).to(dev).train()
with torch.no_grad():                                       # identical start (replicas are built from one seed in practice)
    for t in (vq._codebook.embed, vq._codebook.embed_avg):
        dist.broadcast(t, src=0)
for step in range(3):
    x = torch.randn(4, 10, 64, device=dev) * 0.1 + 5.0      # far from most codes: many die and get re-seeded
    vq(x)
    replicas_identical(vq._codebook, f"dead-code re-seed step {step}")
vqk = tvq.VectorQuantize(64, 16, sync_codebook=True, kmeans_init=True, kmeans_iters=5).to(dev).train()
vqk(torch.randn(4, 50, 64, device=dev))
replicas_identical(vqk._codebook, "k-means init")
assert bool(vqk._codebook.initted.item())

torch.cuda.synchronize()
dist.barrier()
print("DP_OK", rank, flush=True)
import faulthandler
faulthandler.dump_traceback_later(90, exit=True)      # a teardown that hangs leaves a stack, not a stuck GPU box
del vq, vqk
dist.destroy_process_group()
