"""Experiment helper (not part of the product): build libtvq with -DTVQ_PROFILE_PHASES, run the
fused forward at a large N and print CTA 0's per-phase clock totals per epilogue group."""
import ctypes, os, subprocess, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "t-vq-vae-trajgen_b200", "csrc")
SO = os.path.join(CSRC, "_prof", "libtvq_prof.so")

VARIANTS = {"base": []}

def so_path(v):
    return SO.replace(".so", f"_{v}.so")

def build():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    procs = []
    for v, flags in VARIANTS.items():
        procs.append(subprocess.Popen(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                           "-DTVQ_PROFILE_PHASES", *flags, "-shared", "-Xcompiler", "-fPIC", "-o", so_path(v),
                           os.path.join(CSRC, "tvq_api.cu")], cwd=CSRC))
    for pr in procs:
        assert pr.wait() == 0

def run(n=1 << 22, k=32, d=128, train=True, variant='base'):
    lib = ctypes.CDLL(so_path(variant))
    print('variant', variant)
    vp, i64, i, u, f, sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_uint, ctypes.c_float, ctypes.c_size_t
    lib.tvq_forward.argtypes = [vp, vp, i64, i, i, u, f, vp, vp, vp, vp, vp, sz, vp]
    lib.tvq_workspace_bytes.restype = sz
    lib.tvq_workspace_bytes.argtypes = [i64, i, i]
    dev = torch.device("cuda:0")
    x = torch.randn(n, d, device=dev); e = torch.randn(k, d, device=dev)
    idx = torch.empty(n, dtype=torch.int64, device=dev); q = torch.empty_like(x)
    stats = torch.empty(((k + 3) & ~3) + k * d, device=dev); sc = torch.empty(8, device=dev)
    wsb = lib.tvq_workspace_bytes(n, k, d); ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
    flags = (1 if train else 0) | 2
    for _ in range(3):
        rc = lib.tvq_forward(x.data_ptr(), e.data_ptr(), n, k, d, flags, 1.0, idx.data_ptr(), q.data_ptr(), stats.data_ptr(),
                             sc.data_ptr(), ws.data_ptr(), wsb, None)
        assert rc == 0, rc
    torch.cuda.synchronize()
    gt = (ctypes.c_ulonglong * 4)()
    lib.tvq_debug_gt(gt, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lib.tvq_forward(x.data_ptr(), e.data_ptr(), n, k, d, flags, 1.0, idx.data_ptr(), q.data_ptr(), stats.data_ptr(),
                    sc.data_ptr(), ws.data_ptr(), wsb, None)
    e1.record(); torch.cuda.synchronize()
    out = (ctypes.c_ulonglong * 32)()
    lib.tvq_debug_phases(out)
    lib.tvq_debug_gt(gt, 0)
    print(f"  globaltimer: first CTA start -> last CTA end = {(gt[1]-gt[0])/1000:.2f} us; max CTA clocks {gt[2]} ; max clocks to end of main loop {gt[3]}")
    names = ["wait_full", "wait_tmem", "scan", "apply_rest", "release", "ap_load+shfl", "ap_butterfly", "ap_decide+out", "ap_rmw"]
    tiles = (n + 63) // 64
    per_cta = tiles / 148
    print(f"n={n} k={k} d={d} train={train}: {e0.elapsed_time(e1):.3f} ms, ~{per_cta:.0f} tiles/CTA ({per_cta/2:.0f} per group)")
    print("  kernel timeline (clk from start) thread0:", [int(out[8 + i]) for i in range(8)], " thread64:", [int(out[16 + 8 + i]) for i in range(8)])
    for g in range(2):
        tot = sum(out[g * 16 + j] for j in range(8))
        print(f" group {g}: total {tot} clk;", ", ".join(f"{names[j]}={out[g*16+j]/max(1,per_cta/2):.0f}" for j in range(8)), "(clk per tile)")

def run_step(n, k=32, d=128, variant='base'):
    """Fused train step (forward + EMA, one launch): kernel timeline of CTA 0 and the global span."""
    lib = ctypes.CDLL(so_path(variant))
    dev = torch.device("cuda:0")
    lib.tvq_workspace_bytes.restype = ctypes.c_size_t
    lib.tvq_workspace_bytes.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_int]
    vp, i64, i, f, dbl, sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_size_t
    lib.tvq_train_step.argtypes = [vp, vp, vp, vp, vp, i64, i, i, f, dbl, dbl, vp, vp, vp, vp, vp, vp, sz, vp]
    x = torch.randn(n, d, device=dev); e = torch.randn(k, d, device=dev); cs = torch.zeros(k, device=dev); avg = e.clone(); prev = e.clone()
    idx = torch.empty(n, dtype=torch.int64, device=dev); q = torch.empty_like(x); sc = torch.empty(8, device=dev)
    wsb = lib.tvq_workspace_bytes(n, k, d); ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
    call = lambda: lib.tvq_train_step(x.data_ptr(), e.data_ptr(), cs.data_ptr(), avg.data_ptr(), prev.data_ptr(), n, k, d, 1.0, 0.8, 1e-5,
                                      idx.data_ptr(), q.data_ptr(), sc.data_ptr(), None, None, ws.data_ptr(), wsb, None)
    for _ in range(3):
        assert call() == 0
    torch.cuda.synchronize()
    gt = (ctypes.c_ulonglong * 4)()
    lib.tvq_debug_gt(gt, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); call(); e1.record(); torch.cuda.synchronize()
    out = (ctypes.c_ulonglong * 32)()
    lib.tvq_debug_phases(out)
    lib.tvq_debug_gt(gt, 0)
    print(f"train_step n={n}: event {e0.elapsed_time(e1)*1000:.1f} us; first CTA start -> last CTA end {(gt[1]-gt[0])/1000:.2f} us; max CTA clocks {gt[2]}; max clocks to end of main loop {gt[3]}")
    print("  timeline (clk from start: prologue, main loop, sync, dump, flush, loss, finish) thread0:", [int(out[8 + j]) for j in range(8)], " thread64:", [int(out[16 + 8 + j]) for j in range(8)])
    names = ["wait_full", "wait_tmem", "scan", "apply_rest", "release", "ap_load+shfl", "ap_butterfly", "ap_decide+out"]
    for g in range(2):
        print(f"  CTA0 group {g} totals:", ", ".join(f"{names[j]}={int(out[g*16+j])}" for j in range(8)))
    pa = (ctypes.c_ulonglong * 24)()
    lib.tvq_debug_phases_all(pa)
    nm = ["wait_full", "wait_tmem", "scan", "apply_rest", "release", "ap_load", "ap_butterfly", "ap_gather+st+store+codes", "ap_tmem_rmw", "ap_decide"]
    for g in range(2):
        print(f"  CTA0 group {g} all phases:", ", ".join(f"{nm[j]}={int(pa[g*12+j])}" for j in range(10)))
    tl = (ctypes.c_ulonglong * 32)()
    lib.tvq_debug_tiles(tl)
    print("  CTA0 warp2 per tile [landed, scores, scan done, apply done]:", [[int(tl[4 * t + j]) for j in range(4)] for t in range(4)])
    ts = []
    for _ in range(8):
        e0.record(); call(); e1.record(); torch.cuda.synchronize(); ts.append(round(e0.elapsed_time(e1) * 1000, 1))
    print("  8 more eager launches, event us each:", ts)

if __name__ == "__main__":
    if "--step" in sys.argv:
        for n in (18432, 76800, 1 << 20):
            run_step(n)
        sys.exit(0)
    if "--build" in sys.argv:
        build()
    else:
        run(n=18432); run(n=76800); run(n=1<<20)
