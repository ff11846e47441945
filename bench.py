#!/usr/bin/env python
"""bench.py — throughput of the VQ hot path (BASELINE.json metric: VQ latents quantised / second).

Workload (BASELINE.json configs[1]): the VectorQuantize work of ONE stage-1 training step at batch
1024 synthetic trajectories of configs/config.yaml shape — LF codebook 18*1024 = 18 432 latents +
HF codebook 75*1024 = 76 800 latents, K = 32, D = 128, train mode: distance + argmin + gather +
straight-through + commitment loss + EMA statistics + EMA update, then the backward.  One "step"
= both codebooks, forward + backward, through the reference-shaped module (tvq_b200.VectorQuantize).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-sweep]

Output: ONE JSON line (rank 0).  See DESIGN.md section 6 for how every field is obtained.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

B_TRAJ = 1024
TOK_LF, TOK_HF = 18, 75           # tokens per trajectory at L=200, n_fft=4
K_CODES, DIM = 32, 128            # configs/config.yaml: codebook_sizes 32/32, hid_dim 128
LATENTS_PER_STEP = B_TRAJ * (TOK_LF + TOK_HF)
N_INPUT_SETS = 4                  # rotate inputs: 4 x 195 MB of x/g per rank > 126 MB L2
METRIC = "vq_latents_per_sec"
UNIT = "latents/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), float(d.get("bf16_tflops", 1590.0)), "measured"
        except Exception:
            pass
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [t.strip() for t in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------- reference arm

def make_cpu_inputs(seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B_TRAJ, TOK_LF, DIM, generator=g), torch.randn(B_TRAJ, TOK_HF, DIM, generator=g),
            torch.randn(B_TRAJ, TOK_LF, DIM, generator=g), torch.randn(B_TRAJ, TOK_HF, DIM, generator=g))


def cpu_reference_step(states, xl, xh, gl, gh):
    """The reference's CPU torch path for the same step (oracle/vq_oracle.py restates vq.py bitwise)."""
    import vq_oracle as O
    total = 0.0
    for state, x, g in ((states[0], xl, gl), (states[1], xh, gh)):
        xr = x.clone().requires_grad_(True)
        q, ind, loss, ppl = O.vq_forward(state, xr, training=True)
        ((q * g).sum() + loss["loss"].sum()).backward()
        total += float(loss["loss"])
    return total


def run_cpu_baseline(steps, warmup):
    import vq_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    states = [O.new_state(K_CODES, DIM), O.new_state(K_CODES, DIM)]
    inputs = make_cpu_inputs(1)
    for _ in range(warmup):
        cpu_reference_step(states, *inputs)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_step(states, *inputs)
    dt = time.perf_counter() - t0
    return LATENTS_PER_STEP * steps / dt, dt / steps, cores


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, s_per_step, cores = run_cpu_baseline(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} full steps (B=1024: {LATENTS_PER_STEP} latents each), torch "
                                   f"{torch.__version__} CPU fp32, oracle/vq_oracle.py (bitwise restatement of the "
                                   f"reference's vq.py; the Python reference itself cannot travel to the GPU box)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(n_gpus):
    return {"workload": "stage1_vq_train_step_B1024 (BASELINE configs[1]): LF 18432 + HF 76800 latents per GPU, "
                        "K=32, D=128, VectorQuantize train forward (assign+gather+ST+commit loss+EMA) + backward",
            "latents_per_step_per_gpu": LATENTS_PER_STEP, "codebook": [K_CODES, DIM], "batch_trajectories_per_gpu": B_TRAJ,
            "parallelism": f"dp{n_gpus} (batch-sharded; EMA statistics summed over NVLink peer memory inside the forward kernel)",
            "l2_policy": f"inputs rotate over {N_INPUT_SETS} resident batches (> 126 MB L2 between reuses)"}


# -------------------------------------------------------------------------------------- our arm

def our_arm(args):
    import torch.distributed as dist
    import tvq_b200 as tvq

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 kernels have no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    hbm_gbs, bf16_tf, peak_src = load_peaks()

    def note(msg):      # progress on stderr (stdout carries only the JSON line); a hung phase is then visible in the log
        if rank == 0:
            print(f"[bench] {msg}", file=sys.stderr, flush=True)

    torch.manual_seed(0)                                        # identical replicas
    vq_l = tvq.VectorQuantize(DIM, K_CODES, sync_codebook=world > 1).to(dev).train()
    vq_h = tvq.VectorQuantize(DIM, K_CODES, sync_codebook=world > 1).to(dev).train()
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    sets = []
    for _ in range(N_INPUT_SETS):
        sets.append(tuple(torch.randn(B_TRAJ, t, DIM, device=dev, generator=gen).requires_grad_(r)
                          for t, r in ((TOK_LF, True), (TOK_HF, True), (TOK_LF, False), (TOK_HF, False))))

    ones = torch.ones(1, device=dev)

    side_h = torch.cuda.Stream()

    def step(xl, xh, gl, gh):
        """One VQ train step (both codebooks), forward + backward, via the public module API.
        The LF and HF quantisers are independent, so each runs on its own stream (an explicit
        HF-forward-before-LF-forward dependency was tried and is slower: the tail of one forward kernel
        overlaps the head of the other when the hardware is free to schedule them).
        Data-parallel runs do the same: each fused kernel's last CTA waits for that codebook's statistics from every
        peer while the other codebook's kernel proceeds; the data-parallel launch leaves one SM out of its grid, so the
        two kernels can never starve each other of SMs whatever order the ranks start them in (tvq_api.cu)."""
        cur = torch.cuda.current_stream()
        side_h.wait_stream(cur)
        with torch.cuda.stream(side_h):
            qh, ih, lh, ph = vq_h(xh)
            torch.autograd.grad([qh, lh["loss"]], [xh], [gh, ones])
        ql, il, ll, pl = vq_l(xl)
        torch.autograd.grad([ql, ll["loss"]], [xl], [gl, ones])
        cur.wait_stream(side_h)
        return ll["loss"], lh["loss"], il, ih

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, local=False):
        """ms for `steps` calls of fn(i): barrier + synchronize on both sides, max over ranks.  local=True: this rank only
        (sections that rank 0 runs alone must not enter a collective)."""
        sync = torch.cuda.synchronize if local else barrier
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        sync()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1 and not local:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    def graph_timed(fn, reps, local=False):
        """ms per call of fn(i), i = 0..reps-1, replayed from ONE CUDA graph (no host launch gaps); None if capture fails."""
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                kg = torch.cuda.CUDAGraph()
                with torch.cuda.graph(kg, capture_error_mode="thread_local"):
                    for i in range(reps):
                        fn(i)
            torch.cuda.current_stream().wait_stream(side)
            kg.replay()
            return timed(lambda i: kg.replay(), 3, local) / (3 * reps)
        except Exception as exc:
            print(f"[bench] rank {rank}: graph capture failed: {exc}", file=sys.stderr)
            return None

    def clear_grads():
        pass

    # ---- eager path (public API, one Python call per module) -----------------------------------
    def eager(i):
        step(*sets[i % N_INPUT_SETS])
    for i in range(max(args.warmup, 3)):
        eager(i)
    clear_grads()
    eager_ms = timed(eager, args.steps)
    clear_grads()
    note(f"eager: {eager_ms / args.steps * 1e3:.1f} us per step")

    # ---- CUDA-graph replay of the same step (launch-bound regime: 4 small kernels per step).  ONE graph holds
    #      several consecutive training steps (rotating over the resident input batches, EMA state carried from
    #      step to step inside the graph), so the host launch gap is paid once per replay, not once per step.
    graph = None
    graph_ms = None
    steps_per_replay = max(dv for dv in (8, 7, 6, 5, 4, 3, 2, 1) if args.steps % dv == 0)   # EXACTLY args.steps are timed
    n_replays = args.steps // steps_per_replay
    timed_steps = args.steps
    if not args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for s in sets:
                    step(*s)
            torch.cuda.current_stream().wait_stream(side)
            barrier()
            graph = torch.cuda.CUDAGraph()
            # thread_local: the NCCL watchdog thread may touch CUDA while this thread captures
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                for j in range(steps_per_replay):
                    step(*sets[j % N_INPUT_SETS])
        except Exception as exc:   # report, never hide
            print(f"[bench] rank {rank}: CUDA-graph capture failed: {exc}", file=sys.stderr)
            graph = None
        if world > 1:              # every rank must take the same path BEFORE anything is replayed
            torch.cuda.synchronize()
            ok = torch.tensor([1 if graph is not None else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok) == 0:
                graph = None
        if graph is not None:
            for i in range(max(args.warmup, 3)):
                graph.replay()
            graph_ms = timed(lambda i: graph.replay(), n_replays)

    # ---- the timed region that `value` reports, with clocks sampled during it -----------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if graph is not None:
        main_ms = timed(lambda i: graph.replay(), n_replays)
        mode = f"cuda_graph_replay ({steps_per_replay} steps per replay)"
    else:
        timed_steps = args.steps
        main_ms = timed(eager, args.steps)
        mode = "eager"
    clocks = sampler.stop() if rank == 0 else None
    clear_grads()
    value = world * LATENTS_PER_STEP * timed_steps / (main_ms * 1e-3)
    note(f"timed region ({mode}): {main_ms / timed_steps * 1e3:.1f} us per step")

    # ---- e2e: pinned host inputs -> H2D -> step -> D2H of the step's result -------------------------
    #      Double-buffered: the H2D copy of step i+1 runs on a copy stream while step i computes (every step's input is
    #      still copied from pinned host memory inside the timed region; K copies per K timed steps).
    host_sets = [(s[0].detach().cpu().pin_memory(), s[1].detach().cpu().pin_memory()) for s in sets]
    bufs = [(torch.empty_like(sets[0][0]).requires_grad_(True), torch.empty_like(sets[0][1]).requires_grad_(True)) for _ in range(2)]
    copy_s = torch.cuda.Stream()
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]
    for ev in free:
        ev.record()
    h_loss = torch.empty(2, dtype=torch.float32).pin_memory()
    h_il = torch.empty(B_TRAJ, TOK_LF, dtype=torch.int64).pin_memory()
    h_ih = torch.empty(B_TRAJ, TOK_HF, dtype=torch.int64).pin_memory()
    h2d = (bufs[0][0].numel() + bufs[0][1].numel()) * 4
    d2h = 8 + (h_il.numel() + h_ih.numel()) * 8

    def issue_copy(i):
        b = i % 2
        hx_l, hx_h = host_sets[i % N_INPUT_SETS]
        with torch.cuda.stream(copy_s), torch.no_grad():
            copy_s.wait_event(free[b])                     # the step that last used this buffer has finished
            bufs[b][0].copy_(hx_l, non_blocking=True)
            bufs[b][1].copy_(hx_h, non_blocking=True)
            ready[b].record(copy_s)

    def e2e_step(i):
        b = i % 2
        issue_copy(i + 1)                                  # next step's input travels while this step computes
        cur = torch.cuda.current_stream()
        cur.wait_event(ready[b])
        loss_l, loss_h, il, ih = step(bufs[b][0], bufs[b][1], sets[0][2], sets[0][3])
        free[b].record(cur)
        h_loss[0:1].copy_(loss_l.detach(), non_blocking=True)
        h_loss[1:2].copy_(loss_h.detach(), non_blocking=True)
        h_il.copy_(il, non_blocking=True)
        h_ih.copy_(ih, non_blocking=True)
    issue_copy(0)
    for i in range(max(args.warmup, 3)):
        e2e_step(i)
    i0 = max(args.warmup, 3)
    e2e_ms = timed(lambda i: e2e_step(i0 + i), args.steps)
    torch.cuda.synchronize()
    e2e_value = world * LATENTS_PER_STEP * args.steps / (e2e_ms * 1e-3)
    note(f"e2e: {e2e_ms / args.steps * 1e3:.1f} us per step")

    # ---- roofline of the dominant kernel: the HF fused train step (forward + EMA, ONE launch — the kernel the
    #      timed region runs for the HF codebook), timed alone with CUDA events on the launching stream --------
    cb = vq_h._codebook
    ws = cb._workspace(dev)
    flats = [s[1].detach().reshape(-1, DIM) for s in sets]
    n_hf = flats[0].shape[0]
    lib = tvq._lib.load()
    idx = torch.empty(n_hf, dtype=torch.int64, device=dev)
    q = torch.empty_like(flats[0])
    scal = torch.empty(8, device=dev)
    emb, csz, eavg = cb.embed.detach().clone(), cb.cluster_size.detach().clone(), cb.embed_avg.detach().clone()
    st = torch.cuda.current_stream().cuda_stream

    def fwd_kernel(i):
        x = flats[i % N_INPUT_SETS]
        rc = lib.tvq_train_step(x.data_ptr(), emb.data_ptr(), csz.data_ptr(), eavg.data_ptr(), None, n_hf, K_CODES, DIM, 1.0,
                                0.8, 1e-5, idx.data_ptr(), q.data_ptr(), scal.data_ptr(), None, None, ws.buf.data_ptr(),
                                ws.nbytes, st)
        assert rc == 0
    for i in range(5):
        fwd_kernel(i)
    reps = max(args.steps, 20)
    k_ms = timed(fwd_kernel, reps) / reps
    k_mode = "eager launches back to back"
    def fwd_kernel_cur(i):       # same launch on whatever stream is current (graph capture)
        x = flats[i % N_INPUT_SETS]
        rc = lib.tvq_train_step(x.data_ptr(), emb.data_ptr(), csz.data_ptr(), eavg.data_ptr(), None, n_hf, K_CODES, DIM, 1.0,
                                0.8, 1e-5, idx.data_ptr(), q.data_ptr(), scal.data_ptr(), None, None, ws.buf.data_ptr(),
                                ws.nbytes, torch.cuda.current_stream().cuda_stream)
        assert rc == 0
    g_ms = graph_timed(fwd_kernel_cur, reps)
    if g_ms is not None and g_ms < k_ms:
        k_ms, k_mode = g_ms, f"{reps} launches per CUDA-graph replay"
    alg_bytes = n_hf * (8 * DIM + 8)
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    traffic = None
    try:   # DRAM bytes per launch of this kernel from the committed ncu --set full capture (profiles/)
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("fwd_umma_train_hf_n76800")
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s", "frac": achieved / hbm_gbs,
                "traffic": traffic, "peak_source": peak_src,
                "kernel": "fwd_umma_kernel<128,32,train> via tvq_train_step (fused forward + EMA, one launch) on the HF "
                          "codebook, N=76800, " + k_mode,
                "algorithmic_bytes_per_launch": alg_bytes, "us_per_launch": k_ms * 1e3,
                "note": "76 800 latents = 12 us of HBM time: launch/tail-latency regime (SURVEY 7.3-4); `sweep` holds the "
                        "large-N points (BASELINE configs[2]) where the fraction is meaningful"}

    # ---- large-N sweep points (BASELINE configs[2]) ---------------------------------------------------
    sweep = []
    if rank == 0 and not args.no_sweep:
        for (n, k, d) in ((1 << 22, 32, 128), (1 << 22, 512, 64), (1 << 21, 1024, 128), (1 << 20, 4096, 128),
                          (1 << 20, 4096, 256), (1 << 19, 16384, 256)):
            try:
                sweep.append(sweep_point(tvq, dev, n, k, d, hbm_gbs, bf16_tf))
            except Exception as exc:
                sweep.append({"n": n, "k": k, "d": d, "error": str(exc)})

    # ---- the STFT LF/HF front end of the same batch (1024 x 4 x 200), one kernel ----------
    frontend = None
    if rank == 0:
        try:
            xt = [torch.rand(B_TRAJ, 4, 200, device=dev) * 2 - 1 for _ in range(8)]       # 8 x 56 MB of outputs > L2
            outs = [tvq.lf_hf_frontend(x, 4, want=("enc_in_l", "enc_in_h", "x_l", "x_h")) for x in xt]   # warm-up + allocation
            lib_f = tvq._lib.load()

            def fe(i):
                x, o = xt[i % 8], outs[i % 8]
                assert lib_f.tvq_frontend(x.data_ptr(), B_TRAJ, 4, 200, 4, None, o["enc_in_l"].data_ptr(), o["enc_in_h"].data_ptr(),
                                          o["x_l"].data_ptr(), o["x_h"].data_ptr(), torch.cuda.current_stream().cuda_stream) == 0
            fe_ms = graph_timed(fe, 40, local=True) or timed(fe, 40, local=True) / 40
            fe_bytes = B_TRAJ * 4 * (200 * 4 + 2 * (2 * 3 * 201 * 4) + 2 * 200 * 4)
            frontend = {"what": "tvq_frontend: x (1024,4,200) -> enc_in_l, enc_in_h (1024,8,3,201), x_l, x_h (1024,4,200); "
                                "n_fft=4 (stage1.py:101-113, vq_vae.py:179-180)", "us_per_launch": fe_ms * 1e3,
                        "algorithmic_bytes": fe_bytes, "hbm_gbs": fe_bytes / (fe_ms * 1e-3) / 1e9,
                        "frac_of_hbm_peak": fe_bytes / (fe_ms * 1e-3) / 1e9 / hbm_gbs}
        except Exception as exc:
            frontend = {"error": str(exc)}

    # ---- CPU baseline on this host (rank 0, N=1 only) --------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, s_per, cores = run_cpu_baseline(5, 1)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"5 full steps of the same workload ({LATENTS_PER_STEP} latents each, {s_per * 1e3:.1f} ms/step), "
                         "oracle/vq_oracle.py on torch CPU fp32"}

    # per codebook: fused train step (forward + EMA in ONE kernel; data-parallel: its last CTA also sums the
    # statistics of all ranks over NVLink peer memory) + backward
    launches_per_step = 2 * (1 + 1)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": timed_steps, "warmup": max(args.warmup, 3),
            "ms_per_step": main_ms / timed_steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": dict(workload_config(world), timed_mode=mode),
            "trajectories_per_sec": value / (TOK_LF + TOK_HF),
            "eager_ms_per_step": eager_ms / args.steps, "graph_ms_per_step": (graph_ms / timed_steps) if graph_ms else None,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps,
                    "what": "pinned host x (LF+HF) -> H2D (copy stream, double-buffered: step i+1 travels while step i computes) -> VectorQuantize fwd+bwd (eager, public API) -> D2H of loss + indices"},
            "gpu_launches": launches_per_step * timed_steps,
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks, "sweep": sweep, "frontend": frontend,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # the captured graph holds NCCL work: drop it, drain, and leave without the (hanging) teardown
        graph = None
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def sweep_point(tvq, dev, n, k, d, hbm_gbs, bf16_tf):
    """Isolated quantise (train forward incl. EMA statistics) at a BASELINE configs[2] size."""
    g = torch.Generator(device=dev).manual_seed(1)
    reps_in = max(2, int(2.6e8 // (n * d * 4)) + 1)            # distinct inputs totalling > 2 x L2
    xs = [torch.randn(n, d, device=dev, generator=g) for _ in range(min(reps_in, 3))]
    e = torch.randn(k, d, device=dev, generator=g)
    ws = tvq.Workspace(k, d, dev)
    for i in range(3):
        tvq.vq_forward_raw(xs[i % len(xs)], e, ws, train=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for i in range(reps):
        tvq.vq_forward_raw(xs[i % len(xs)], e, ws, train=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    by = n * (8 * d + 8) / (ms * 1e-3) / 1e9
    fl = 2.0 * n * k * d / (ms * 1e-3) / 1e12
    hb, tc = by / hbm_gbs, fl / bf16_tf
    # the bound is the SLOWER of the two rooflines for this shape; frac = that roofline's time / ours
    t_hbm, t_tc = n * (8 * d + 8) / (hbm_gbs * 1e9), 2.0 * n * k * d / (bf16_tf * 1e12)
    return {"n": n, "k": k, "d": d, "mode": "train_forward", "ms": ms, "latents_per_sec": n / (ms * 1e-3),
            "hbm_gbs": by, "tflops": fl, "bound": "hbm" if t_hbm >= t_tc else "tensor", "frac": max(hb, tc),
            "path": "tcgen05 tf32, resident codebook" if k <= 32 and d <= 128 else "tcgen05 bf16 nomination, streamed codebook",
            "peak": "measured copy GB/s / measured cuBLAS bf16 TFLOP/s (MEASURED_PEAKS.json)"}


def main():
    import faulthandler
    faulthandler.dump_traceback_later(240, exit=False, file=sys.stderr)     # a hang leaves a Python stack in the log
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        our_arm(args)


if __name__ == "__main__":
    main()
