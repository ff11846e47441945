"""world_size-2 gloo test (CPU) of the data-parallel host logic: every rank packs its local
statistics in the TVQ layout [counts (padded to 4) | embed_sum (K-major)], the module's ONE
all-reduce hook sums them, and the EMA update applied to the reduced buffer leaves the replicas
bit-identical and equal to the full-batch update (vq.py:229,234 semantics).  The per-shard
arithmetic is the oracle's here (no GPU in this container); the kernels' version of the same
property is tests/test_parity_gpu.py::test_sync_codebook_virtual_ranks."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import tvq_b200 as tvq
    import vq_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        k, d = 16, 32
        torch.manual_seed(7)                                     # identical replicas
        vq = tvq.VectorQuantize(d, k, sync_codebook=True)
        cb = vq._codebook
        assert cb._ddp_active()
        full = torch.randn(2, 6, 40, d, generator=torch.Generator().manual_seed(9))
        ref_state = {n: getattr(cb, n).clone() for n in ("initted", "cluster_size", "embed_avg", "embed")}
        off = tvq.stats_offset(k)
        for step in range(2):
            x = full[step, rank * 3:(rank + 1) * 3].reshape(-1, d)
            ind = O.neg_sq_dist(x, cb.embed).argmax(-1)
            onehot = torch.nn.functional.one_hot(ind, k).float()
            stats = torch.zeros(tvq.stats_len(k, d))
            stats[:k] = onehot.sum(0)
            stats[off:] = (onehot.t() @ x).reshape(-1)           # K-major embed_sum
            cb._all_reduce_stats(stats)                          # the module's hook: ONE packed all-reduce
            # EMA update from the packed buffer (the arithmetic tvq_ema_update performs on the GPU)
            cb.cluster_size.mul_(cb.decay).add_(stats[:k], alpha=1 - cb.decay)
            cb.embed_avg.mul_(cb.decay).add_(stats[off:].view(k, d), alpha=1 - cb.decay)
            n = cb.cluster_size.sum()
            sm = (cb.cluster_size + cb.eps) / (n + k * cb.eps) * n
            cb.embed.copy_(cb.embed_avg / sm.unsqueeze(1))
            # full-batch oracle on the same data
            O.vq_forward(ref_state, full[step].reshape(1, -1, d), training=True)
        gathered = [torch.empty_like(cb.embed) for _ in range(world)]
        dist.all_gather(gathered, cb.embed)
        out.put((rank, bool(torch.equal(gathered[0], gathered[1])),
                 float((cb.embed - ref_state["embed"]).abs().max()), float(ref_state["embed"].abs().max()),
                 float((cb.cluster_size - ref_state["cluster_size"]).abs().max())))
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_two_rank_packed_statistics_allreduce():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = 29650 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get() for _ in range(2)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, identical, err, scale, cerr in results:
        assert identical, "replicas diverged"
        assert err <= 1e-5 * scale, f"rank {rank}: sharded EMA differs from the full-batch update by {err}"
        assert cerr <= 1e-5


def _sampling_worker(rank, world, port, out):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import tvq_b200 as tvq
    from tvq_b200.vq import sample_rows
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)                            # different generators AND different batches per rank
        data = torch.randn(50 + 10 * rank, 8)
        rows_many = sample_rows(data, 16, True)                  # randperm branch (n >= num)
        rows_few = sample_rows(data[:5], 16, True)               # randint branch (n < num)
        local = sample_rows(data, 16, False)                     # single-process behaviour: a rank-local draw
        g = [torch.empty_like(rows_many) for _ in range(world)]
        dist.all_gather(g, rows_many)
        g2 = [torch.empty_like(rows_few) for _ in range(world)]
        dist.all_gather(g2, rows_few)
        # rank 0's rows must come from rank 0's own batch
        from_rank0 = True
        if rank == 0:
            from_rank0 = all(bool((data == r).all(1).any()) for r in rows_many)
        # the peer-exchange decision on a gloo group: agreed "not used", no collective entered, all_reduce path kept
        vq = tvq.VectorQuantize(8, 16, sync_codebook=True)
        px = vq._codebook.setup_data_parallel(torch.device("cpu"))
        out.put((rank, bool(torch.equal(g[0], g[1])), bool(torch.equal(g2[0], g2[1])), from_rank0,
                 bool((local == rows_many).all()), px is None and vq._codebook._px is False))
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_two_rank_sampling_is_replica_consistent():
    """k-means seeds / dead-code replacements (vq.py:67-75) under data parallelism: every rank ends with RANK 0's rows
    (the reference draws rank-locally: replicas would diverge, SURVEY section 8 e); without DDP the draw stays local.
    And the peer-exchange set-up on a non-NCCL group: every rank agrees on the all_reduce path."""
    import warnings
    warnings.filterwarnings("ignore")
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = 29850 + (os.getpid() % 100)
    procs = [ctx.Process(target=_sampling_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get() for _ in range(2)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, same_many, same_few, from0, local_equals, px_off in results:
        assert same_many and same_few, "ranks ended with different rows"
        assert from0, "rank 0's rows are not rows of rank 0's batch"
        assert px_off, "a gloo group must take the all_reduce path, agreed on by all ranks"
        if rank == 1:
            assert not local_equals, "rank 1's local draw should differ from rank 0's broadcast rows"


def test_sync_flag_is_inert_without_a_process_group():
    import tvq_b200 as tvq
    vq = tvq.VectorQuantize(32, 16, sync_codebook=True)
    assert not vq._codebook._ddp_active()
    t = torch.ones(4)
    vq._codebook._all_reduce_stats(t)                            # no group: must be a no-op, not an error
    assert torch.equal(t, torch.ones(4))
