"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

TEST INFRASTRUCTURE ONLY.  Needs /root/reference (absent on the GPU box), so it
is run by hand here and its outputs are committed:

    python oracle/gen_golden.py            # rewrites tests/golden/

The reference ships no tests or golden vectors, so every
fixture is an input/output pair of `timevqvae/models/vq.py` (loaded through the
real package import, with the absent third-party modules stubbed) and, for the
config-1 cases, of `Stage1`'s encoders + `utils/train_utils.py::quantize`.
Each case stores the inputs, the module state before the call, the outputs and
the state after, plus the smallest top-2 score margin (in ulps) so the tests
know that the stored indices are decidable by any fp32 implementation.
"""
from __future__ import annotations

import hashlib
import importlib.abc
import importlib.machinery
import os
import sys
import types
import warnings

import numpy as np
import torch
import torch.nn as nn

warnings.filterwarnings("ignore")
REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


class _Any:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Any()

    def __getattr__(self, n):
        return _Any()


class _StubModule(types.ModuleType):
    def __getattr__(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Any


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Third-party modules the reference imports but this image lacks."""
    ROOTS = ("mlflow", "lightning", "matplotlib", "traffic", "x_transformers", "altair", "cartes",
             "cartopy", "numba", "seaborn", "bluesky", "mpl_toolkits", "openap", "pyproj",
             "shapely", "geopy")

    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in self.ROOTS:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, m):
        if m.__name__ == "lightning":
            m.LightningModule = nn.Module


def load_reference():
    sys.meta_path.insert(0, _StubFinder())
    sys.path.insert(0, REF)
    import timevqvae.models.vq as ref_vq
    import timevqvae.trainers.stage1 as ref_stage1
    import timevqvae.utils.train_utils as ref_tu
    return ref_vq, ref_stage1, ref_tu


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.detach().contiguous().numpy().tobytes()).hexdigest()


def state_of(vq, prefix):
    cb = vq._codebook
    out = {f"{prefix}initted": cb.initted.clone().numpy(), f"{prefix}cluster_size": cb.cluster_size.clone().numpy(),
           f"{prefix}embed_avg": cb.embed_avg.clone().numpy(), f"{prefix}embed": cb.embed.detach().clone().numpy()}
    return out


def margin_ulps(ref_vq, x, embed):
    """Smallest top-2 gap of the reference's own fp32 scores, in ulps."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from vq_oracle import neg_sq_dist, top2_margin_ulps
    flat = x.reshape(-1, x.shape[-1])
    return float(top2_margin_ulps(neg_sq_dist(flat, embed)).min())


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    conv = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        conv[k] = np.asarray(v)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **conv)
    print(f"{name:28s} {os.path.getsize(path) / 1024:8.1f} KiB")


def loss_arrays(vq_loss, prefix="out_"):
    commit = vq_loss["commit_loss"]
    return {prefix + "loss": vq_loss["loss"].detach().numpy(),
            prefix + "commit_loss": np.float32(commit.item() if torch.is_tensor(commit) else commit),
            prefix + "commit_is_tensor": np.bool_(torch.is_tensor(commit))}


def main():
    ref_vq, ref_stage1, ref_tu = load_reference()
    VQ = ref_vq.VectorQuantize

    # -- 1. the reference's only known-answer block, vq.py:410-424 ------------------------------
    torch.manual_seed(0)
    x = torch.rand((1024, 32, 128))
    vq = VQ(dim=128, codebook_size=512)
    pre = state_of(vq, "pre_")
    q, ind, loss, ppl = vq(x)
    assert ind[0, 0].item() == 87, "vq.py:421 known answer"
    m = margin_ulps(ref_vq, x, torch.from_numpy(pre["pre_embed"]))
    save("smoke_main", x_sha=np.array(sha(x)), embed_sha=np.array(sha(torch.from_numpy(pre["pre_embed"]))),
         out_ind=ind.numpy().astype(np.int16), out_q_sha=np.array(sha(q)), out_q_first=q[:2].detach(),
         out_perplexity=ppl, min_margin_ulps=m, **loss_arrays(loss),
         post_cluster_size=vq._codebook.cluster_size, post_embed_avg=vq._codebook.embed_avg,
         post_embed=vq._codebook.embed)

    # -- 2. config 1: Stage1 encoders (configs/config.yaml) -> quantize() glue, B = 32 ----------
    cfg = ref_tu.load_yaml_param_settings(os.path.join(REF, "configs", "config.yaml"))
    torch.manual_seed(0)
    np.random.seed(0)
    stage1 = ref_stage1.Stage1(200, 4, cfg)
    stage1.train()
    xb = torch.rand(32, 4, 200) * 2 - 1
    for tag, enc, vqm in (("lf", stage1.encoder_l, stage1.vq_model_l), ("hf", stage1.encoder_h, stage1.vq_model_h)):
        z = enc(xb).detach()
        pre = state_of(vqm, "pre_")
        zr = z.clone().requires_grad_(True)
        zq, s, vq_loss, ppl = ref_tu.quantize(zr, vqm)
        gq = torch.randn(zq.shape, generator=torch.Generator().manual_seed(7))
        ((zq * gq).sum() + vq_loss["loss"].sum()).backward()
        b, c, h, w = z.shape
        m = margin_ulps(ref_vq, z.permute(0, 2, 3, 1).reshape(b, h * w, c), torch.from_numpy(pre["pre_embed"]))
        post = state_of(vqm, "post_")
        # same module, eval mode (tokenise path, models/maskgit.py:117-134); buffers must not move
        vqm.eval()
        with torch.no_grad():
            zq_e, s_e, vq_loss_e, ppl_e = ref_tu.quantize(z, vqm)
        assert all(np.array_equal(v, state_of(vqm, "post_")[k]) for k, v in post.items())
        vqm.train()
        # the HF tensors are 1.2 MB each: keep every 8th trajectory of the big outputs, hash the rest
        keep = slice(None) if tag == "lf" else slice(0, None, 8)
        save(f"cfg1_{tag}", z=z, g_zq_seed=7, keep_step=(1 if tag == "lf" else 8),
             out_zq_kept=zq[keep], out_zq_sha=np.array(sha(zq)), out_ind=s.numpy().astype(np.int16),
             out_perplexity=ppl, out_grad_z_kept=zr.grad[keep], min_margin_ulps=m,
             eval_ind=s_e.numpy().astype(np.int16), eval_perplexity=ppl_e, eval_zq_sha=np.array(sha(zq_e)),
             eval_zq_kept=zq_e[keep], **loss_arrays(vq_loss_e, "eval_"),
             **pre, **loss_arrays(vq_loss), **post)

    # -- 3. three consecutive training steps (EMA sequencing) ----------------------------------
    torch.manual_seed(11)
    vq = VQ(dim=32, codebook_size=16, decay=0.8, commitment_weight=0.25)
    arrays = dict(state_of(vq, "pre_"))
    margins = []
    for step in range(3):
        x = torch.randn(4, 50, 32) * (1.0 + 0.5 * step)
        margins.append(margin_ulps(ref_vq, x, vq._codebook.embed.clone()))
        q, ind, loss, ppl = vq(x)
        arrays.update({f"x{step}": x, f"out{step}_q": q, f"out{step}_ind": ind.numpy().astype(np.int16),
                       f"out{step}_perplexity": ppl, **loss_arrays(loss, f"out{step}_"),
                       **state_of(vq, f"post{step}_")})
    save("train_3steps", min_margin_ulps=min(margins), commitment_weight=np.float32(0.25), **arrays)

    # -- 4. layout / head / projection variants --------------------------------------------------
    torch.manual_seed(12)
    vq = VQ(dim=64, codebook_size=24, heads=2, codebook_dim=32)
    x = torch.randn(3, 20, 64)
    pre = state_of(vq, "pre_")
    q, ind, loss, ppl = vq(x)
    save("heads2_train", x=x, out_q=q, out_ind=ind.numpy().astype(np.int16), out_perplexity=ppl, **pre,
         **loss_arrays(loss), **state_of(vq, "post_"))

    torch.manual_seed(13)
    vq = VQ(dim=128, codebook_size=32, codebook_dim=64)          # SURVEY section 8 note 1: the "32x64" variant
    x = torch.randn(2, 30, 128, requires_grad=True)
    pre = state_of(vq, "pre_")
    q, ind, loss, ppl = vq(x)
    gq = torch.randn(q.shape, generator=torch.Generator().manual_seed(8))
    ((q * gq).sum() + loss["loss"].sum()).backward()
    save("proj64_train", x=x, g_q=gq, out_q=q, out_ind=ind.numpy().astype(np.int16), out_perplexity=ppl,
         out_grad_x=x.grad, w_in=vq.project_in.weight, b_in=vq.project_in.bias, w_out=vq.project_out.weight,
         b_out=vq.project_out.bias, grad_w_in=vq.project_in.weight.grad, grad_b_in=vq.project_in.bias.grad,
         grad_w_out=vq.project_out.weight.grad, grad_b_out=vq.project_out.bias.grad,
         **pre, **loss_arrays(loss), **state_of(vq, "post_"))

    torch.manual_seed(14)
    vq = VQ(dim=32, codebook_size=16, accept_image_fmap=True)
    x = torch.randn(2, 32, 3, 5)
    pre = state_of(vq, "pre_")
    q, ind, loss, ppl = vq(x)
    save("image_fmap_train", x=x, out_q=q, out_ind=ind.numpy().astype(np.int16), out_perplexity=ppl, **pre,
         **loss_arrays(loss), **state_of(vq, "post_"))

    torch.manual_seed(15)
    vq = VQ(dim=32, codebook_size=16, channel_last=False)
    x = torch.randn(2, 32, 17)
    pre = state_of(vq, "pre_")
    q, ind, loss, ppl = vq(x)
    save("channel_first_train", x=x, out_q=q, out_ind=ind.numpy().astype(np.int16), out_perplexity=ppl, **pre,
         **loss_arrays(loss), **state_of(vq, "post_"))

    # -- 5. RNG-consuming branches: dead-code expiry, k-means init, stochastic sampling ---------
    torch.manual_seed(16)
    vq = VQ(dim=16, codebook_size=32, threshold_ema_dead_code=2)
    x = torch.randn(1, 40, 16)                       # N=40 >= K -> randperm branch; most codes are dead
    pre = state_of(vq, "pre_")
    torch.manual_seed(100)
    q, ind, loss, ppl = vq(x)
    save("dead_code_randperm", x=x, rng_seed=100, out_q=q, out_ind=ind.numpy().astype(np.int16),
         out_perplexity=ppl, **pre, **loss_arrays(loss), **state_of(vq, "post_"))

    torch.manual_seed(17)
    vq = VQ(dim=16, codebook_size=32, threshold_ema_dead_code=2)
    x = torch.randn(1, 10, 16)                       # N=10 < K -> randint branch
    pre = state_of(vq, "pre_")
    torch.manual_seed(101)
    q, ind, loss, ppl = vq(x)
    save("dead_code_randint", x=x, rng_seed=101, out_q=q, out_ind=ind.numpy().astype(np.int16),
         out_perplexity=ppl, **pre, **loss_arrays(loss), **state_of(vq, "post_"))

    torch.manual_seed(18)
    vq = VQ(dim=16, codebook_size=8, kmeans_init=True, kmeans_iters=10)
    x = torch.randn(2, 100, 16)
    pre = state_of(vq, "pre_")
    torch.manual_seed(102)
    q, ind, loss, ppl = vq(x)
    save("kmeans_init_train", x=x, rng_seed=102, out_q=q, out_ind=ind.numpy().astype(np.int16),
         out_perplexity=ppl, **pre, **loss_arrays(loss), **state_of(vq, "post_"))

    torch.manual_seed(19)
    vq = VQ(dim=32, codebook_size=32)
    vq.eval()
    x = torch.randn(4, 60, 32)
    pre = state_of(vq, "pre_")
    torch.manual_seed(103)
    with torch.no_grad():
        q, ind, loss, ppl = vq(x, 0.5)               # svq_temp (trainers/stage3.py:119-124)
    save("svq_temp_eval", x=x, rng_seed=103, svq_temp=np.float32(0.5), out_q=q,
         out_ind=ind.numpy().astype(np.int16), out_perplexity=ppl, **pre, **loss_arrays(loss),
         **state_of(vq, "post_"))

    # -- 6. token decode hand-off (models/maskgit.py:465-470) -----------------------------------
    from einops import rearrange
    torch.manual_seed(20)
    embed = torch.randn(32, 128)
    s = torch.randint(0, 32, (5, 75))
    zq = torch.nn.functional.embedding(s, embed)
    zq = rearrange(zq, "b n c -> b c n")
    zq = rearrange(zq, "b c (h w) -> b c h w", h=3, w=25)
    save("decode_gather", tokens=s.numpy().astype(np.int16), embed=embed, h=3, w=25, out_zq=zq)

    # -- 7. sync_codebook=True on two gloo ranks (vq.py:155,229,234) -----------------------------
    import torch.multiprocessing as mp
    mp.set_start_method("spawn", force=True)
    q_out = mp.get_context("spawn").SimpleQueue()
    procs = [mp.get_context("spawn").Process(target=_ddp_worker, args=(r, 2, q_out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q_out.get() for _ in range(2)]
    for p in procs:
        p.join()
    arrays = {}
    for r in results:
        arrays.update(r)
    save("sync_codebook_2rank", **arrays)


def _ddp_worker(rank, world, q_out):
    warnings.filterwarnings("ignore")
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = "29611"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ref_vq, _, _ = load_reference()
    torch.manual_seed(21)                                     # same init on both ranks
    vq = ref_vq.VectorQuantize(dim=32, codebook_size=16, sync_codebook=True)
    out = {}
    if rank == 0:
        out.update({k: v for k, v in state_of(vq, "pre_").items()})
    xs = torch.randn(2, 3, 40, 32, generator=torch.Generator().manual_seed(22))   # (step, rank*..)
    full = torch.randn(2, 6, 40, 32, generator=torch.Generator().manual_seed(23))  # (step, global batch)
    for step in range(2):
        x = full[step, rank * 3:(rank + 1) * 3]
        q, ind, loss, ppl = vq(x)
        out.update({f"r{rank}_x{step}": x.numpy(), f"r{rank}_out{step}_q": q.detach().numpy(),
                    f"r{rank}_out{step}_ind": ind.numpy().astype(np.int16),
                    f"r{rank}_out{step}_perplexity": ppl.numpy(),
                    f"r{rank}_out{step}_loss": loss["loss"].detach().numpy(),
                    **{k: v for k, v in state_of(vq, f"r{rank}_post{step}_").items()}})
    q_out.put(out)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
