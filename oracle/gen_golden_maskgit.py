"""Generate tests/golden/maskgit_*.npz from the UNMODIFIED reference (timevqvae/models/maskgit.py) in this container:
the loop body of first_pass (:300-346) is executed line by line with the reference's own mask_by_random_topk and
torch's own Categorical / uniform_ draws from a seeded generator; the noise those draws consumed is recorded by
re-seeding and drawing again in the same order.  TEST INFRASTRUCTURE ONLY."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_golden as G


def reference_step(mg_cls, logits, s, mask_id, mask_len, temperature, seed):
    torch.manual_seed(seed)
    sampled_ids = torch.distributions.categorical.Categorical(logits=logits).sample()
    unknown_map = s == mask_id
    sampled_ids = torch.where(unknown_map, sampled_ids, s)
    probs = F.softmax(logits, dim=-1)
    selected_probs = torch.gather(probs, dim=-1, index=sampled_ids.unsqueeze(-1)).squeeze()
    selected_probs = torch.where(unknown_map, selected_probs, torch.Tensor([torch.inf]))
    ml = torch.full((logits.shape[0], 1), float(mask_len))
    masking = mg_cls.mask_by_random_topk(None, ml, selected_probs, temperature=temperature, device="cpu")
    s_new = torch.where(masking, mask_id, sampled_ids)
    return s_new, sampled_ids, masking


def main():
    G.load_reference()
    from timevqvae.models.maskgit import MaskGIT
    for name, (b, n, k, mask_len, temp, frac_known, seed) in {
            "maskgit_lf_t0": (8, 18, 32, 15, 4.0 * (1 - 1 / 10), 0.0, 1), "maskgit_lf_mid": (8, 18, 32, 7, 4.0 * (1 - 6 / 10), 0.5, 2),
            "maskgit_hf": (6, 75, 32, 30, 2.0, 0.4, 3), "maskgit_k512": (3, 40, 512, 11, 1.0, 0.25, 4),
            "maskgit_last": (4, 18, 32, 0, 0.0, 0.8, 5)}.items():
        g = torch.Generator().manual_seed(100 + seed)
        logits = torch.randn(b, n, k, generator=g) * 2.5
        mask_id = k
        s = torch.randint(0, k, (b, n), generator=g)
        # the reference keeps exactly (previous mask_len) unknown tokens per row: choose them at random
        n_unknown = max(mask_len, int(round(n * (1 - frac_known))))
        for r in range(b):
            perm = torch.randperm(n, generator=g)[:n_unknown]
            s[r, perm] = mask_id
        s_new, sampled, masking = reference_step(MaskGIT, logits, s, mask_id, mask_len, temp, seed)
        torch.manual_seed(seed)                                   # the noise those calls consumed, in the same order
        q = torch.empty(b * n, k).exponential_(1).reshape(b, n, k)
        u = torch.zeros(b, n).uniform_(0, 1)
        np.savez_compressed(os.path.join(G.OUT, name + ".npz"), logits=logits.numpy(), s=s.numpy(), mask_id=np.int64(mask_id),
                            mask_len=np.int64(mask_len), temperature=np.float64(temp), q=q.numpy(), u=u.numpy(),
                            s_new=s_new.numpy(), sampled=sampled.numpy(), masking=masking.numpy())
        print(name, int(masking.sum()), "masked of", b * n)


if __name__ == "__main__":
    main()
