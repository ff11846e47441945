"""Data-parallel EMA update through the fused NVLink peer-memory kernel (tvq_ema_update_dp).

world = 1 runs on any B200 (the kernel pushes to, waits on and reduces its own buffer) and must equal
tvq_ema_update bit for bit.  The real exchange needs at least two GPUs: the test launches tests/dp_worker.py with one
process per GPU under torchrun (NCCL process group for the rendezvous, 127.0.0.1).  The worker drives the SHIPPED fused
kernels (tvq_train_step_dp, the channels-first tvq_train_step_qcf with peers, tvq_ema_update_dp) against the reference's
own 2-rank gloo run (tests/golden/sync_codebook_2rank.npz) and against a full-batch single-GPU run at BASELINE configs[3]
shapes, and asserts bit-identical replicas after every step; skipped on a one-GPU box (logs of the 2- and 8-GPU runs of
the same worker are kept under profiles/)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def tvq():
    import tvq_b200
    assert torch.cuda.is_available()
    return tvq_b200


def test_world1_equals_ema_update(tvq):
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    k, d = 32, 128
    lib = tvq._lib.load()
    stats = torch.zeros(tvq.stats_len(k, d), device=dev)
    stats[:k] = torch.randint(0, 50, (k,), device=dev).float()
    stats[tvq.stats_offset(k):] = torch.randn(k * d, device=dev)
    base = [torch.rand(k, device=dev) * 5, torch.randn(k, d, device=dev), torch.randn(k, d, device=dev)]
    ws = tvq.Workspace(k, d, dev)
    a = [t.clone() for t in base]
    prev_a = torch.empty(k, d, device=dev)
    tvq.vq_ema_update(stats, a[0], a[1], a[2], prev_a, 0.8, 1e-5, ws)
    buf = torch.zeros(int(lib.tvq_exchange_bytes(k, d, 1)), dtype=torch.uint8, device=dev)
    peers = torch.tensor([buf.data_ptr()], dtype=torch.int64, device=dev)
    b = [t.clone() for t in base]
    prev_b = torch.empty(k, d, device=dev)
    for step in range(3):       # three steps: both parities and the in-buffer step counter
        b = [t.clone() for t in base]
        rc = lib.tvq_ema_update_dp(stats.data_ptr(), peers.data_ptr(), 0, 1, b[0].data_ptr(), b[1].data_ptr(), b[2].data_ptr(),
                                   prev_b.data_ptr(), k, d, 0.8, 1e-5, torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        torch.cuda.synchronize()
        for x, y in zip(a, b):
            assert torch.equal(x, y)
        assert torch.equal(prev_a, prev_b)


def run_worker(world, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dp_worker.py"), ROOT]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=420)
    assert out.returncode == 0 and out.stdout.count("DP_OK") == world, out.stdout[-3000:] + out.stderr[-6000:]
    return out.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_peer_exchange():
    """Two ranks: includes the golden vectors of the reference's own 2-rank run."""
    out = run_worker(2, 29533)
    assert "golden sync_codebook_2rank reproduced" in out


@pytest.mark.skipif(torch.cuda.device_count() < 4, reason="needs four or more GPUs")
def test_all_ranks_peer_exchange():
    run_worker(torch.cuda.device_count(), 29534)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_module_on_a_device_that_is_not_current(tvq):
    """ONE process, two GPUs: a module living on cuda:1 while cuda:0 is the current device must launch on cuda:1 (the C ABI
    works on the current device: the wrappers switch around every call) and the per-device kernel-attribute caches must
    configure the > 48 KB shared-memory kernels on BOTH devices.  Same seeds -> the same results."""
    torch.cuda.set_device(0)
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        torch.manual_seed(3)
        vq = tvq.VectorQuantize(128, 32).to(dev).train()
        big = tvq.VectorQuantize(64, 512).to(dev).train()             # streamed-codebook kernel (its own attribute cache)
        g = torch.Generator().manual_seed(4)
        x = torch.randn(8, 75, 128, generator=g).to(dev).requires_grad_(True)
        q, ind, loss, ppl = vq(x)
        (q.sum() + loss["loss"].sum()).backward()
        z = torch.randn(8, 128, 3, 25, generator=g).to(dev)
        zq, ind2, _, _ = tvq.quantize(z, vq)
        qb, indb, _, _ = big(torch.randn(4, 300, 64, generator=g).to(dev))
        assert torch.cuda.current_device() == 0
        torch.cuda.synchronize(dev)
        outs.append([t.detach().cpu() for t in (q, ind, x.grad, qb, indb, loss["loss"], zq, vq._codebook.embed, big._codebook.embed)])
    exact, close = 5, 4          # the first call of each module is bit-exact; what follows an EMA update (atomics) is 1e-5
    for i, (a, b) in enumerate(zip(*outs)):
        if i < exact:
            assert torch.equal(a, b), i
        else:
            torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-5 * float(b.abs().max()))
