"""Experiment helper (not part of the product): build libtvq with -DTVQ_STREAM_PROF and print CTA 0's per-role
wait / work clock totals of the streamed-codebook forward.  usage: profile_stream.py build | run n k d [train]"""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "t-vq-vae-trajgen_b200", "csrc")
SO = os.path.join(CSRC, "_prof", "libtvq_sprof.so")

VARIANTS = {"": []}
_OLD_VARIANTS = {"": [], "noslow": ["-DTVQ_ABL_NOSLOW"], "noe2": ["-DTVQ_ABL_NOE2"], "noslow_nold": ["-DTVQ_ABL_NOSLOW", "-DTVQ_ABL_NOLD"], "noslow_nomma": ["-DTVQ_ABL_NOSLOW", "-DTVQ_ABL_NOMMA"],
            "noslow_noe2": ["-DTVQ_ABL_NOSLOW", "-DTVQ_ABL_NOE2"]}

def so_of(v):
    return SO if not v else SO.replace(".so", f"_{v}.so")

def build():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    procs = [subprocess.Popen(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-DTVQ_STREAM_PROF",
                               *fl, "-shared", "-Xcompiler", "-fPIC", "-o", so_of(v), os.path.join(CSRC, "tvq_api.cu")], cwd=CSRC)
             for v, fl in VARIANTS.items()]
    for pr in procs:
        assert pr.wait() == 0

def run(n, k, d, train, variant=""):
    import torch
    lib = ctypes.CDLL(so_of(variant))
    print('variant', variant or 'base')
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
    for _ in range(2):
        rc = lib.tvq_forward(x.data_ptr(), e.data_ptr(), n, k, d, flags, 1.0, idx.data_ptr(), q.data_ptr(), stats.data_ptr(),
                             sc.data_ptr(), ws.data_ptr(), wsb, None)
        assert rc == 0, rc
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lib.tvq_forward(x.data_ptr(), e.data_ptr(), n, k, d, flags, 1.0, idx.data_ptr(), q.data_ptr(), stats.data_ptr(),
                    sc.data_ptr(), ws.data_ptr(), wsb, None)
    e1.record(); torch.cuda.synchronize()
    out = (ctypes.c_ulonglong * 128)()
    lib.tvq_debug_stream_prof(out)
    tiles = -(-n // 128); per_cta = -(-tiles // 148)
    nt = 128 if d > 128 else 256
    nct = -(-k // nt)
    print(f"n={n} k={k} d={d} train={train}: {e0.elapsed_time(e1):.3f} ms; CTA0: {per_cta} row tiles x {nct} code tiles")
    def show(name, base, labels):
        vals = [int(out[base + j]) for j in range(8)]
        print(f"  {name:9s}", ", ".join(f"{l}={v/per_cta/1000:.1f}k" for l, v in zip(labels, vals) if l), "(clk per row tile)")
    show("producer", 0, ["wait_e_empty", "wait_b_empty", "", "", "", "", "", "total"])
    show("mma", 8, ["wait_a_full", "wait_t_empty", "wait_b_full", "", "", "", "", "total"])
    show("scan w2", 16, ["wait_rows", "wait_e2", "wait_t_full", "scan", "merge", "apply", "apply:fetch+issue", "apply:2nd pass"])
    show("convert", 24, ["wait_a_empty", "wait_r_empty", "work", "issue_blk", "wait_x", "fence+arrive", "", ""])
    for w in range(0):
        show(f"scan w{w+2}", 64 + 8 * w, ["wait_rows", "wait_e2", "wait_t_full", "scan", "merge", "apply", "", ""])

if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    else:
        run(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), bool(int(sys.argv[5])) if len(sys.argv) > 5 else False,
            sys.argv[6] if len(sys.argv) > 6 else "")
