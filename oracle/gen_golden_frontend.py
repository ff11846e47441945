"""Generate tests/golden/frontend_*.npz by running the UNMODIFIED reference front end in this container
(/root/reference/timevqvae/utils/train_utils.py, trainers/stage1.py:101-113).  TEST INFRASTRUCTURE ONLY."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_golden as G

OUT = G.OUT


def main():
    _, _, tu = G.load_reference()
    for name, (b, c, l, n_fft) in {"frontend_cfg1": (6, 4, 200, 4), "frontend_nfft8": (3, 2, 96, 8), "frontend_odd": (2, 3, 101, 4), "frontend_interp": (2, 5, 333, 8)}.items():
        g = torch.Generator().manual_seed(7 + l)
        x = torch.rand(b, c, l, generator=g) * 2 - 1
        xf = tu.time_to_timefreq(x, n_fft, c)
        u_l, u_h = tu.zero_pad_high_freq(xf), tu.zero_pad_low_freq(xf)
        x_l = F.interpolate(tu.timefreq_to_time(u_l, n_fft, c), l, mode="linear")
        x_h = F.interpolate(tu.timefreq_to_time(u_h, n_fft, c), l, mode="linear")
        np.savez_compressed(os.path.join(OUT, name + ".npz"), x=x.numpy(), n_fft=np.int64(n_fft), xf=xf.numpy(),
                            enc_in_l=tu.zero_pad_high_freq(xf, copy=True).numpy(), enc_in_h=tu.zero_pad_low_freq(xf, copy=True).numpy(),
                            x_l=x_l.numpy(), x_h=x_h.numpy())
        print(name, tuple(xf.shape), tuple(x_l.shape))


def decoder_side():
    """models/vq_vae.py:259-262: pad_func -> timefreq_to_time -> interpolate, forward and autograd backward."""
    _, _, tu = G.load_reference()
    for name, (b, c, l, n_fft) in {"istft_cfg1": (4, 4, 200, 4), "istft_nfft8": (2, 3, 96, 8), "istft_interp": (2, 2, 333, 8)}.items():
        g = torch.Generator().manual_seed(11 + l)
        hop = n_fft // 4
        u = torch.randn(b, 2 * c, n_fft // 2 + 1, l // hop + 1, generator=g)
        gy = torch.randn(b, c, l, generator=g)
        rec = {"u": u.numpy(), "g_y": gy.numpy(), "n_fft": np.int64(n_fft)}
        for band, pad in (("all", lambda t: t), ("lf", tu.zero_pad_high_freq), ("hf", tu.zero_pad_low_freq)):
            uu = u.clone().requires_grad_(True)
            y = F.interpolate(tu.timefreq_to_time(pad(uu), n_fft, c), l, mode="linear")
            (y * gy).sum().backward()
            rec["y_" + band] = y.detach().numpy()
            rec["g_u_" + band] = uu.grad.numpy()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        print(name, tuple(u.shape))


if __name__ == "__main__":
    decoder_side()
    main()
