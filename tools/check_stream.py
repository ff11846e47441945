"""Bring-up check of the streamed-codebook tcgen05 path: parity against the CUDA-core path and the C
oracle on small shapes, then timings at BASELINE configs[2] sizes.  Run on a B200 (gpurun)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import tvq_b200 as tvq
import vq_canon as C
C.build()
DEV = "cuda"
ok = True
shapes = [(128, 64, 64), (130, 100, 128), (1000, 256, 64), (777, 33, 64), (5000, 100, 100), (600, 16, 256), (3000, 512, 64),
          (2049, 1000, 128), (1500, 70, 256), (4096, 2500, 32), (20000, 4096, 128), (9000, 777, 252), (300000, 512, 64)]
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    shapes = shapes[:3]
for (n, k, d) in shapes:
    for train in (True, False):
        torch.manual_seed(n + k + d)
        x = (torch.randn(n, d) * 1.3 + 0.2).to(DEV)
        e = torch.randn(k, d).to(DEV)
        ws = tvq.Workspace(k, d, torch.device(DEV))
        off = tvq.stats_offset(k)
        idx_u, q_u, sc_u = tvq.vq_forward_raw(x, e, ws, train=train)
        st_u = ws.stats.clone()
        idx_s, q_s, sc_s = tvq.vq_forward_raw(x, e, ws, train=train, flags=tvq._lib.F_NO_UMMA)
        st_s = ws.stats.clone()
        torch.cuda.synchronize()
        bad = int((idx_u != idx_s).sum())
        qeq = torch.equal(q_u, q_s)
        ceq = torch.equal(st_u[:k], st_s[:k])
        resc = sc_u.view(torch.int32)[4:6].tolist()
        es = float((st_u[off:] - st_s[off:]).abs().max() / (st_s[off:].abs().max() + 1e-30)) if train else 0.0
        ls = float((sc_u[0] - sc_s[0]).abs() / (sc_s[0].abs() + 1e-30)) if train else 0.0
        pp = float((sc_u[1] - sc_s[1]).abs() / sc_s[1].abs())
        cok = True
        if n <= 20000:
            cok = np.array_equal(idx_u.cpu().numpy(), C.assign(x.cpu().numpy(), e.cpu().numpy()))
        good = bad == 0 and qeq and ceq and es < 1e-5 and ls < 1e-5 and pp < 1e-5 and cok
        ok &= good
        print(f"n={n} k={k} d={d} train={train}: idx_mismatch={bad} q_eq={qeq} counts_eq={ceq} esum_rel={es:.2e} loss_rel={ls:.2e} "
              f"ppl_rel={pp:.2e} canon={cok} rescored={resc} {'OK' if good else 'FAIL'}", flush=True)
# near-ties / duplicated codes: candidate lists overflow -> spill buffer -> exhaustive scan; first index must win
for (n, k, d, dup, noise) in [(3000, 640, 64, 20, 0.0), (2000, 1024, 128, 6, 0.0), (4000, 2048, 128, 12, 1e-3), (1500, 512, 256, 40, 0.0)]:
    torch.manual_seed(n + k)
    k = (k // dup) * dup
    base = torch.randn(k // dup, d)
    e = (base.repeat_interleave(dup, 0) + noise * torch.randn(k, d)).to(DEV)
    x = torch.randn(n, d).to(DEV)
    ws = tvq.Workspace(k, d, torch.device(DEV))
    idx_u, q_u, sc_u = tvq.vq_forward_raw(x, e, ws, train=True)
    idx_s, q_s, sc_s = tvq.vq_forward_raw(x, e, ws, train=True, flags=tvq._lib.F_NO_UMMA)
    torch.cuda.synchronize()
    bad = int((idx_u != idx_s).sum())
    cok = np.array_equal(idx_u.cpu().numpy(), C.assign(x.cpu().numpy(), e.cpu().numpy()))
    good = bad == 0 and cok and torch.equal(q_u, q_s)
    ok &= good
    print(f"dup n={n} k={k} d={d} dup={dup} noise={noise}: idx_mismatch={bad} canon={cok} rescored={sc_u.view(torch.int32)[4:6].tolist()} {'OK' if good else 'FAIL'}", flush=True)
print("PARITY", "OK" if ok else "FAIL", flush=True)
if not ok or (len(sys.argv) > 1 and sys.argv[1] == "quick"):
    sys.exit(0 if ok else 1)
# ---- timings
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
res = []
for (n, k, d) in [(1 << 22, 512, 64), (1 << 22, 512, 128), (1 << 21, 1024, 128), (1 << 20, 4096, 64), (1 << 20, 4096, 128), (1 << 20, 4096, 256),
                  (1 << 19, 16384, 128), (1 << 19, 16384, 256)]:
    g = torch.Generator(device=DEV).manual_seed(1)
    xs = [torch.randn(n, d, device=DEV, generator=g) for _ in range(2)]
    e = torch.randn(k, d, device=DEV, generator=g)
    ws = tvq.Workspace(k, d, torch.device(DEV))
    for train in (False, True):
        for i in range(2):
            tvq.vq_forward_raw(xs[i % 2], e, ws, train=train)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for i in range(reps):
            _, _, sc = tvq.vq_forward_raw(xs[i % 2], e, ws, train=train)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        by = n * (8 * d + 8) / (ms * 1e-3) / 1e9
        fl = 2.0 * n * k * d / (ms * 1e-3) / 1e12
        r = dict(n=n, k=k, d=d, train=train, ms=round(ms, 4), glat_s=round(n / ms / 1e6, 3), hbm_frac=round(by / peaks["hbm_gbs"], 3),
                 tflops=round(fl, 1), tc_frac=round(fl / peaks["bf16_tflops"], 3), rescored=sc.view(torch.int32)[4:6].tolist())
        res.append(r)
        print(json.dumps(r), flush=True)
