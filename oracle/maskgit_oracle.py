"""CPU restatement (torch) of ONE MaskGIT decoding iteration after the transformer — TEST INFRASTRUCTURE ONLY.

Follows /root/reference/timevqvae/models/maskgit.py:
  first_pass / second_pass loop body   :300-346 / :364-410  (sample ids, keep known tokens, softmax, confidence of the
                                                             sampled ids, inf for known tokens, re-mask the least confident)
  mask_by_random_topk                  :238-267             (log(p + 1e-5) + temperature * Gumbel noise, top-k smallest)
The reference draws its noise inside torch: Categorical(logits).sample() is torch.multinomial with one sample, i.e.
argmax(probs / q) with q = empty_like(probs).exponential_(1) [probed in this container], and the Gumbel noise is
-log(-log(zeros_like(p).uniform_(0, 1))).  Here both noise tensors are INPUTS (q, u), drawn by the caller in that order,
so that the same generator state gives the reference's result bit for bit (tests/golden/maskgit_*.npz).
"""
import torch
import torch.nn.functional as F


def maskgit_step(logits, s, mask_token_id, mask_len, temperature, q, u):
    """logits (b,n,K) fp32, s (b,n) int64, q (b,n,K) ~ Exp(1), u (b,n) ~ U(0,1) -> (s_new, sampled_ids, masking)."""
    b, n, k = logits.shape
    probs = F.softmax(logits, dim=-1)
    sampled = (probs.reshape(-1, k) / q.reshape(-1, k)).argmax(-1).reshape(b, n)          # maskgit.py:305-307
    unknown = s == mask_token_id                                                         # :310-312
    sampled = torch.where(unknown, sampled, s)                                           # :313-315
    sel = torch.gather(probs, -1, sampled.unsqueeze(-1)).squeeze(-1)                      # :322-324
    sel = torch.where(unknown, sel, torch.tensor(float("inf")))                          # :325-328
    eps = 1e-20
    gumbel = -torch.log((-torch.log(u.clamp(min=eps))).clamp(min=eps))                   # :245-253
    conf = torch.log(sel + 1e-5) + temperature * gumbel                                  # :255-257
    masking = torch.zeros(b, n, dtype=torch.bool)
    if mask_len > 0:
        ind = torch.topk(conf, k=int(mask_len), dim=-1, largest=False).indices            # :259-261
        masking.scatter_(1, ind, True)                                                   # :262-265
    s_new = torch.where(masking, torch.tensor(mask_token_id), sampled)                   # :346
    return s_new, sampled, masking
