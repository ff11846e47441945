#!/usr/bin/env python
"""bench.py — throughput of the VQ hot path and of the stage-1 step around it (BASELINE.json metric:
"VQ latents quantized/sec and stage-1 trajectories/sec at 1/2/4/8 B200 vs CPU").

Main line (`metric`/`value`): BASELINE configs[1] — the VectorQuantize work of ONE stage-1 training step at batch 1024
synthetic trajectories of configs/config.yaml shape: LF 18*1024 = 18 432 + HF 75*1024 = 76 800 latents, K = 32, D = 128, train
mode (distance + argmin + gather + straight-through + commitment loss + EMA statistics + EMA update) and its backward, both
codebooks, through the reference-shaped module (tvq_b200.VectorQuantize).  `value` = latents / s with inputs resident in
HBM; `e2e` = the same through the public API from pinned host buffers.

Extra keys (each measured in this run):
  stage1        the whole stage-1 optimisation step (encoders -> quantize() -> decoders -> losses -> backward -> AdamW) of the
                harness in t-vq-vae-trajgen_b200/stage1.py: trajectories / s, weak (1024 per GPU) and strong (1024 / N per GPU)
                scaling, end to end from pinned host trajectories, the VQ kernels' share of the step, the unmodified reference
                on the host CPU beside it                                                   (configs[1], configs[3])
  dp_parity     N > 1: the data-parallel step checked in THIS run — replicas bit-identical, exchanged counts == world * n,
                one step against oracle/vq_oracle.py on the gathered batch; a failure exits non-zero
  parity        indices of the timed batch against the torch fp32 oracle, every row, with the count of un-decidable rows
  sweep         configs[2]: K in {512..16384} x D in {64,128,256} x N in {2^20, 2^22 (, 2^24)}, eval-assign / train-forward /
                forward+backward, each against its roofline, CPU per-latent baseline per (K, D)
  generation    configs[4]: 10 000 trajectories in batches of 32, 10 + 1 MaskGIT iterations (maskgit_step kernel on synthetic
                logits), decode_tokens, both decoders
  config0       configs[0]: the B = 32 stage-1 forward (and the codebook_dim = 64 variant of SURVEY section 8 Note 1)

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-sweep] [--quick]

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


# ------------------------------------------------------------------------------- reference arm (host CPU)

def make_cpu_inputs(seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B_TRAJ, TOK_LF, DIM, generator=g), torch.randn(B_TRAJ, TOK_HF, DIM, generator=g),
            torch.randn(B_TRAJ, TOK_LF, DIM, generator=g), torch.randn(B_TRAJ, TOK_HF, DIM, generator=g))


def cpu_vq_arm():
    """(kind, step_fn): the reference's own VectorQuantize on the host CPU for the timed workload — the UNMODIFIED
    timevqvae/models/vq.py from oracle/_ref when present ("reference"), else oracle/vq_oracle.py, its bitwise
    restatement ("port")."""
    try:
        import ref_loader
        ref_vq = ref_loader.load()[0]
        torch.manual_seed(0)
        vqs = [ref_vq.VectorQuantize(DIM, K_CODES).train(), ref_vq.VectorQuantize(DIM, K_CODES).train()]

        def step(xl, xh, gl, gh):
            for vq, x, g in ((vqs[0], xl, gl), (vqs[1], xh, gh)):
                xr = x.clone().requires_grad_(True)
                q, ind, loss, ppl = vq(xr)
                ((q * g).sum() + loss["loss"].sum()).backward()
        return "reference", step, "timevqvae/models/vq.py (unmodified, oracle/_ref)"
    except Exception:
        import vq_oracle as O
        torch.manual_seed(0)
        states = [O.new_state(K_CODES, DIM), O.new_state(K_CODES, DIM)]

        def step(xl, xh, gl, gh):
            for state, x, g in ((states[0], xl, gl), (states[1], xh, gh)):
                xr = x.clone().requires_grad_(True)
                q, ind, loss, ppl = O.vq_forward(state, xr, training=True)
                ((q * g).sum() + loss["loss"].sum()).backward()
        return "port", step, "oracle/vq_oracle.py (bitwise restatement of the reference's vq.py)"


def run_cpu_vq(steps, warmup):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kind, step, what = cpu_vq_arm()
    inputs = make_cpu_inputs(1)
    for _ in range(warmup):
        step(*inputs)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(*inputs)
    dt = time.perf_counter() - t0
    return {"value": LATENTS_PER_STEP * steps / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{steps} full steps of the same workload (B=1024: {LATENTS_PER_STEP} latents each, {dt / steps * 1e3:.1f} ms/step), "
                      f"torch {torch.__version__} CPU fp32, {what}"}, dt / steps


def run_cpu_stage1(batch=256, steps=2):
    """The UNMODIFIED reference Stage1 (trainers/stage1.py) on the host CPU: forward + backward + AdamW, train mode."""
    try:
        import numpy as np
        import yaml
        import ref_loader
        _, ref_s1, _, root = ref_loader.load()
        cfg = yaml.safe_load(open(os.path.join(root, "configs", "config.yaml")))
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        torch.manual_seed(0); np.random.seed(0)
        model = ref_s1.Stage1(200, 4, cfg).train()
        opt = torch.optim.AdamW(model.parameters(), lr=cfg["exp_params"]["lr"])
        x = torch.rand(batch, 4, 200) * 2 - 1
        y = torch.zeros(batch, 1, dtype=torch.long)

        def step():
            r, v, p = model((x, y), batch_idx=1)
            loss = r["LF.time"] + r["HF.time"] + v["LF"]["loss"] + v["HF"]["loss"]
            opt.zero_grad(); loss.backward(); opt.step()
        step()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        dt = (time.perf_counter() - t0) / steps
        return {"traj_per_sec": batch / dt, "cores": cores, "kind": "reference",
                "sample": f"{steps} steps at batch {batch} ({dt * 1e3:.0f} ms/step), unmodified trainers/stage1.py + AdamW, torch CPU fp32"}
    except Exception as exc:
        return {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cpu, s_per_step = run_cpu_vq(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "stage1": run_cpu_stage1(),
    }
    print(json.dumps(line))


def workload_config(n_gpus):
    return {"workload": "stage1_vq_train_step_B1024 (BASELINE configs[1]): LF 18432 + HF 76800 latents per GPU, "
                        "K=32, D=128, VectorQuantize train forward (assign+gather+ST+commit loss+EMA) + backward",
            "latents_per_step_per_gpu": LATENTS_PER_STEP, "codebook": [K_CODES, DIM], "batch_trajectories_per_gpu": B_TRAJ,
            "parallelism": f"dp{n_gpus} (batch-sharded; EMA statistics summed over NVLink peer memory inside the forward kernel)",
            "l2_policy": f"inputs rotate over {N_INPUT_SETS} resident batches (> 126 MB L2 between reuses)"}


# -------------------------------------------------------------------------------------- our arm

class Ctx:
    """Per-process timing helpers (device, distributed world, barriers, event timing)."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the B200 kernels have no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
                os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the one JSON line
            dist.init_process_group("nccl", device_id=self.dev)

    def note(self, msg):      # progress on stderr (stdout carries only the JSON line)
        if self.rank == 0:
            print(f"[bench] {msg}", file=sys.stderr, flush=True)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps, local=False):
        """ms for `steps` calls of fn(i): barrier + synchronize on both sides, CUDA events, max over ranks.  local=True: this
        rank only (sections that rank 0 runs alone must not enter a collective)."""
        sync = torch.cuda.synchronize if local else self.barrier
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        sync()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1 and not local:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms)

    def graph_timed(self, fn, reps, local=False, replays=3):
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
            return self.timed(lambda i: kg.replay(), replays, local) / (replays * reps)
        except Exception as exc:
            print(f"[bench] rank {self.rank}: graph capture failed: {exc}", file=sys.stderr)
            return None


def claim_stdout():
    """stdout carries exactly ONE line, the JSON: everything else any library prints to file descriptor 1 (NCCL's version
    banner, for one) is sent to stderr; returns the writer for the JSON line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def our_arm(args):
    import tvq_b200 as tvq
    json_out = claim_stdout()
    c = Ctx()
    dist, world, rank, dev, note = c.dist, c.world, c.rank, c.dev, c.note
    hbm_gbs, bf16_tf, peak_src = load_peaks()
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = True                       # the reference CLI's setting (scripts/train.py:19): affects the
    torch.set_float32_matmul_precision("high")                   # harness convolutions only; the VQ kernels decide in fp32/fp64

    torch.manual_seed(0)                                        # identical replicas
    defer = args.defer if world > 1 else 0
    vq_l = tvq.VectorQuantize(DIM, K_CODES, sync_codebook=world > 1, defer_exchange=defer).to(dev).train()
    vq_h = tvq.VectorQuantize(DIM, K_CODES, sync_codebook=world > 1, defer_exchange=defer).to(dev).train()
    if not args.no_sm_split:
        # the two quantisers run on two streams: share the SMs out in proportion to their latents so that the two persistent
        # forward launches (one whole SM's shared memory per CTA) are resident together instead of one after the other
        vq_l._codebook.sm_share, vq_h._codebook.sm_share = 0.25, 0.75       # measured best (tools/time_split.py): 72 -> 65 us
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    sets = []
    for _ in range(N_INPUT_SETS):
        sets.append(tuple(torch.randn(B_TRAJ, t, DIM, device=dev, generator=gen).requires_grad_(r)
                          for t, r in ((TOK_LF, True), (TOK_HF, True), (TOK_LF, False), (TOK_HF, False))))
    ones = torch.ones(1, device=dev)
    side_h = torch.cuda.Stream()

    def step(xl, xh, gl, gh):
        """One VQ train step (both codebooks), forward + backward, via the public module API.  The LF and HF quantisers are
        independent, so each runs on its own stream.  Data-parallel runs do the same: each fused kernel's last CTA waits for
        that codebook's statistics from every peer while the other codebook's kernel proceeds; the data-parallel launch
        leaves one SM out of its grid, so the two kernels can never starve each other of SMs (tvq_api.cu)."""
        cur = torch.cuda.current_stream()
        side_h.wait_stream(cur)
        with torch.cuda.stream(side_h):
            qh, ih, lh, ph = vq_h(xh)
            torch.autograd.grad([qh, lh["loss"]], [xh], [gh, ones])
            vq_h._codebook.join_pending()                 # (deferred exchange: the finalize kernel rejoins its stream)
        ql, il, ll, pl = vq_l(xl)
        torch.autograd.grad([ql, ll["loss"]], [xl], [gl, ones])
        vq_l._codebook.join_pending()
        cur.wait_stream(side_h)
        return ll["loss"], lh["loss"], il, ih

    # ---- eager path (public API, one Python call per module) -----------------------------------
    def eager(i):
        step(*sets[i % N_INPUT_SETS])
    for i in range(max(args.warmup, 3)):
        eager(i)
    eager_ms = c.timed(eager, args.steps)
    note(f"eager: {eager_ms / args.steps * 1e3:.1f} us per step")

    # ---- CUDA-graph replay of the same step (launch-bound regime: 4 small kernels per step).  ONE graph holds several
    #      consecutive training steps (rotating over the resident input batches, EMA state carried from step to step
    #      inside the graph), so the host launch gap is paid once per replay, not once per step.
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
            c.barrier()
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
            graph_ms = c.timed(lambda i: graph.replay(), n_replays)

    # ---- the timed region that `value` reports, with clocks sampled during it -----------------------
    sampler = ClockSampler(c.local)
    if rank == 0:
        sampler.start()
    if graph is not None:
        main_ms = c.timed(lambda i: graph.replay(), n_replays)
        mode = f"cuda_graph_replay ({steps_per_replay} steps per replay)"
    else:
        main_ms = c.timed(eager, args.steps)
        mode = "eager"
    clocks = sampler.stop() if rank == 0 else None
    value = world * LATENTS_PER_STEP * timed_steps / (main_ms * 1e-3)
    vq_us_per_step = main_ms / timed_steps * 1e3
    note(f"timed region ({mode}): {vq_us_per_step:.1f} us per step")

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
    e2e_ms = c.timed(lambda i: e2e_step(i0 + i), args.steps)
    torch.cuda.synchronize()
    e2e_value = world * LATENTS_PER_STEP * args.steps / (e2e_ms * 1e-3)
    note(f"e2e: {e2e_ms / args.steps * 1e3:.1f} us per step ({h2d / (e2e_ms / args.steps * 1e-3) / 1e9:.1f} GB/s H2D per rank)")
    del host_sets, bufs

    # ---- data-parallel self-check of THIS run (the driver's test box has one GPU) --------------------------------------
    dp_parity = None
    if world > 1:
        dp_parity = dp_parity_check(c, tvq)
        note(f"dp_parity: {dp_parity}")

    # ---- parity of the timed batch against the torch fp32 oracle, every row (rank 0; CPU work, ~0.2 s) ----------------
    parity = None
    if rank == 0:
        parity = parity_check(tvq, dev, sets[0])
        note(f"parity: {parity}")

    # ---- roofline of the dominant kernel: the HF fused train step (forward + EMA, ONE launch — the kernel the
    #      timed region runs for the HF codebook), timed alone with CUDA events on the launching stream --------
    roofline = dominant_kernel_roofline(c, tvq, vq_h, sets, hbm_gbs, peak_src, max(args.steps, 20))

    # ---- stage-1 step (BASELINE metric (ii); configs[1] and configs[3]) --------------------------------------------------
    stage1 = None
    if not args.no_stage1:
        try:
            stage1 = stage1_bench(c, tvq, args, vq_us_per_step)
        except Exception as exc:
            stage1 = {"error": f"{type(exc).__name__}: {exc}"[:300]}
            if world > 1:
                raise
        note(f"stage1: {json.dumps(stage1)[:400]}")

    # ---- single-GPU sections: configs[2] sweep, configs[4] generation, configs[0], the STFT front end -----------------
    sweep, generation, config0, frontend = [], None, None, None
    if rank == 0 and world == 1:
        if not args.no_sweep:
            sweep = sweep_all(tvq, dev, hbm_gbs, bf16_tf, args.quick, note)
        try:
            generation = generation_bench(c, tvq, args.quick)
        except Exception as exc:
            generation = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        note(f"generation: {json.dumps(generation)[:300]}")
        try:
            config0 = config0_bench(c, tvq, args)
        except Exception as exc:
            config0 = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        frontend = frontend_bench(c, tvq, hbm_gbs)

    # ---- CPU baseline on this host (rank 0, N=1 only) --------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu, _ = run_cpu_vq(5, 1)

    # per codebook: fused train step (forward + EMA in ONE kernel; data-parallel: its last CTA also sums the
    # statistics of all ranks over NVLink peer memory) + backward
    launches_per_step = 2 * (1 + 1)
    failed = bool(dp_parity and not dp_parity.get("ok", False))
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": timed_steps, "warmup": max(args.warmup, 3),
            "ms_per_step": main_ms / timed_steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": dict(workload_config(world), timed_mode=mode),
            "eager_ms_per_step": eager_ms / args.steps, "graph_ms_per_step": (graph_ms / timed_steps) if graph_ms else None,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps, "h2d_gbs_per_rank": h2d / (e2e_ms / args.steps * 1e-3) / 1e9,
                    "what": "pinned host x (LF+HF) -> H2D (copy stream, double-buffered: step i+1 travels while step i computes) -> VectorQuantize fwd+bwd (eager, public API) -> D2H of loss + indices"},
            "gpu_launches": launches_per_step * timed_steps,
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "stage1_traj_per_sec": (stage1 or {}).get("weak", {}).get("traj_per_sec") if isinstance(stage1, dict) else None,
            "stage1": stage1, "dp_parity": dp_parity, "parity": parity,
            "parity_rows_undecidable": parity.get("rows_undecidable") if parity else None,
            "sweep": sweep, "generation": generation, "config0": config0, "frontend": frontend,
        }
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    teardown(c, graph, failed)


def teardown(c, graph, failed):
    """Leave cleanly: drop the captured graphs, drain, destroy the process group.  (Round 1 left multi-rank runs through
    os._exit because the interpreter's own teardown hung: CUDA graphs and symmetric-memory buffers were still alive when
    the NCCL communicator was torn down at exit.  Releasing them FIRST and destroying the group explicitly is the fix; a
    watchdog still turns a teardown that does not finish into an exit instead of a stuck GPU box.)"""
    del graph
    torch.cuda.synchronize()
    if c.world > 1:
        import faulthandler
        import gc
        import threading
        gc.collect()
        c.dist.barrier()
        sys.stdout.flush(); sys.stderr.flush()

        def bail():
            faulthandler.dump_traceback(file=sys.stderr)
            os._exit(3 if failed else 0)
        t = threading.Timer(60.0, bail)
        t.daemon = True
        t.start()
        c.dist.destroy_process_group()
        t.cancel()
    if failed:
        sys.exit(3)


def dp_parity_check(c, tvq):
    """N > 1, after the timed region: ONE data-parallel step of fresh replicas on the configs[1] shapes, checked three ways —
    (a) replicas bit-identical (all_gather of embed / embed_avg / cluster_size), (b) the exchanged counts sum to
    world * n (cluster_size after one step from zero = counts * (1 - decay)), (c) indices / buffers against
    oracle/vq_oracle.py run on the GATHERED batch (rank 0, host CPU).  The kernels under test are the ones the timed region
    ran (tvq_train_step_dp with the peer exchange)."""
    import vq_oracle as O
    dist, world, rank, dev = c.dist, c.world, c.rank, c.dev
    res = {"ok": True, "replicas_identical": True, "idx_mismatch": 0, "rows_undecidable": 0, "max_rel": 0.0, "counts_total_ok": True,
           "fused_peer_exchange": True}
    for toks in (TOK_LF, TOK_HF):
        torch.manual_seed(5 + toks)
        vq = tvq.VectorQuantize(DIM, K_CODES, sync_codebook=True).to(dev).train()
        pre = {k: getattr(vq._codebook, k).detach().cpu().clone() for k in ("initted", "cluster_size", "embed_avg", "embed")}
        n_loc = 128 * toks
        x = torch.randn(128, toks, DIM, device=dev, generator=torch.Generator(device=dev).manual_seed(900 + rank + toks))
        q, ind, loss, ppl = vq(x)
        res["fused_peer_exchange"] &= bool(vq._codebook._px)
        cb = vq._codebook
        for name in ("cluster_size", "embed_avg", "embed"):
            a = getattr(cb, name).detach().contiguous()
            g = [torch.empty_like(a) for _ in range(world)]
            dist.all_gather(g, a)
            res["replicas_identical"] &= all(torch.equal(g[0], t) for t in g)
        counts = torch.round(cb.cluster_size.double() / (1.0 - cb.decay))
        res["counts_total_ok"] &= float(counts.sum()) == float(world * n_loc)
        gx = [torch.empty_like(x) for _ in range(world)]
        gi = [torch.empty_like(ind) for _ in range(world)]
        dist.all_gather(gx, x.contiguous())
        dist.all_gather(gi, ind.contiguous())
        if rank == 0:
            xa = torch.cat(gx).cpu()
            state = {k: v.clone() for k, v in pre.items()}
            q_ref, ind_ref, loss_ref, ppl_ref = O.vq_forward(state, xa, training=True)      # full batch == all-reduced statistics
            got = torch.cat(gi).cpu().reshape(-1)
            bad = torch.nonzero(got != ind_ref.reshape(-1)).reshape(-1)
            res["idx_mismatch"] += int(bad.numel())
            if bad.numel():
                flat = xa.reshape(-1, DIM)
                m = O.top2_margin_ulps(O.neg_sq_dist(flat[bad], pre["embed"]))
                res["rows_undecidable"] += int((m <= 2).sum())
            for name in ("cluster_size", "embed_avg", "embed"):
                ref = state[name]
                rel = float((getattr(cb, name).detach().cpu() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
                res["max_rel"] = max(res["max_rel"], rel)
        cb.check_peer_errors()
    flag = torch.tensor([1 if (res["replicas_identical"] and res["counts_total_ok"] and res["idx_mismatch"] == res["rows_undecidable"]
                               and res["max_rel"] <= 1e-5 and res["fused_peer_exchange"]) else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    res["ok"] = bool(int(flag))
    res["what"] = ("one data-parallel step of fresh replicas (128 trajectories per rank, LF and HF shapes): replicas bit-identical, "
                   "counts sum == world * n, indices exact (mismatches only on rows the fp32 reference cannot decide) and buffers "
                   "within max_rel of oracle/vq_oracle.py on the gathered batch")
    return res


def parity_check(tvq, dev, s):
    """Indices of the timed batch (LF 18 432 + HF 76 800 rows) against the reference's formula on the host CPU (torch fp32,
    oracle/vq_oracle.py::assign_chunked), EVERY row; a mismatch counts as un-decidable when the reference's own two best
    scores are within 2 ulps."""
    import vq_oracle as O
    torch.manual_seed(77)
    e = torch.randn(K_CODES, DIM)
    out = {"rows": 0, "idx_mismatch": 0, "rows_undecidable": 0}
    for x in (s[0], s[1]):
        flat = x.detach().reshape(-1, DIM).contiguous()
        ws = tvq.Workspace(K_CODES, DIM, dev)
        idx, _, _ = tvq.vq_forward_raw(flat, e.to(dev), ws, train=True)
        xc = flat.cpu()
        ref = O.assign_chunked(xc, e)
        bad = torch.nonzero(idx.cpu() != ref).reshape(-1)
        out["rows"] += flat.shape[0]
        out["idx_mismatch"] += int(bad.numel())
        if bad.numel():
            out["rows_undecidable"] += int((O.top2_margin_ulps(O.neg_sq_dist(xc[bad], e)) <= 2).sum())
    out["ok"] = out["idx_mismatch"] == out["rows_undecidable"]
    return out


def dominant_kernel_roofline(c, tvq, vq_h, sets, hbm_gbs, peak_src, reps):
    dev = c.dev
    cb = vq_h._codebook
    ws = cb._workspace(dev)
    flats = [s[1].detach().reshape(-1, DIM) for s in sets]
    n_hf = flats[0].shape[0]
    lib = tvq._lib.load()
    idx = torch.empty(n_hf, dtype=torch.int64, device=dev)
    q = torch.empty_like(flats[0])
    scal = torch.empty(8, device=dev)
    emb, csz, eavg = cb.embed.detach().clone(), cb.cluster_size.detach().clone(), cb.embed_avg.detach().clone()

    def fwd_kernel(i):       # the launch on whatever stream is current (eager timing and graph capture)
        x = flats[i % N_INPUT_SETS]
        rc = lib.tvq_train_step(x.data_ptr(), emb.data_ptr(), csz.data_ptr(), eavg.data_ptr(), None, n_hf, K_CODES, DIM, 1.0,
                                0.8, 1e-5, idx.data_ptr(), q.data_ptr(), scal.data_ptr(), None, None, ws.buf.data_ptr(),
                                ws.nbytes, torch.cuda.current_stream().cuda_stream)
        assert rc == 0
    for i in range(5):
        fwd_kernel(i)
    k_ms = c.timed(fwd_kernel, reps) / reps
    k_mode = "eager launches back to back"
    g_ms = c.graph_timed(fwd_kernel, reps)
    if g_ms is not None and g_ms < k_ms:
        k_ms, k_mode = g_ms, f"{reps} launches per CUDA-graph replay"
    alg_bytes = n_hf * (8 * DIM + 8)
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    traffic = None
    try:   # DRAM bytes per launch of this kernel from the committed ncu --set full capture (profiles/)
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("fwd_umma_train_hf_n76800")
    except Exception:
        pass
    return {"bound": "hbm", "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s", "frac": achieved / hbm_gbs,
            "traffic": traffic, "peak_source": peak_src,
            "kernel": "fwd_umma_kernel<128,32,train> via tvq_train_step (fused forward + EMA, one launch) on the HF "
                      "codebook, N=76800, " + k_mode,
            "algorithmic_bytes_per_launch": alg_bytes, "us_per_launch": k_ms * 1e3,
            "note": "76 800 latents = 12 us of HBM time: launch/tail-latency regime (SURVEY 7.3-4); `sweep` holds the "
                    "large-N points (BASELINE configs[2]) where the fraction is meaningful"}


# ------------------------------------------------------------------------------------------------ stage 1

def stage1_bench(c, tvq, args, vq_us_per_step):
    """Whole stage-1 optimisation steps through the harness (stage1.py): one CUDA-graph replay per step.
    weak: 1024 trajectories per GPU; strong (N > 1): 1024 / N per GPU.  e2e: each step's trajectories come from pinned host
    memory (H2D of 3.3 MB inside the timed region) and the loss is read back."""
    import numpy as np
    dist, world, rank, dev = c.dist, c.world, c.rank, c.dev
    steps = max(10, min(args.steps, 30))
    out = {"what": "Stage1 harness step: STFT front end (1 kernel) -> 2 x [conv encoder -> quantize() -> conv decoder -> band ISTFT] "
                   "-> MSE(LF) + L1(HF) + VQ losses -> backward -> gradient all-reduce (N > 1) -> AdamW; torch convolutions "
                   "(TF32, as the reference CLI), the VQ / front-end / ISTFT kernels of this repo, one CUDA-graph replay per step",
           "steps": steps}

    def one(batch, key):
        torch.manual_seed(0); np.random.seed(0)
        cfg = tvq.stage1.default_config()
        cfg["VQ-VAE"]["sync_codebook"] = world > 1
        cfg["VQ-VAE"]["defer_exchange"] = args.defer if world > 1 else 0
        model = tvq.Stage1(200, 4, cfg).to(dev).to(memory_format=torch.channels_last)    # NHWC convolutions: 30 -> 20 ms per step
        tr = tvq.Stage1Trainer(model, (batch, 4, 200), use_graph=not args.no_graph)
        tr.warmup_and_capture(3)
        g = torch.Generator(device=dev).manual_seed(300 + rank)
        xs = [torch.rand(batch, 4, 200, device=dev, generator=g) * 2 - 1 for _ in range(4)]
        for i in range(3):
            tr.step(xs[i % 4])
        ms = c.timed(lambda i: tr.step(xs[i % 4]), steps) / steps
        # e2e: pinned host trajectories in, loss out, every step
        hx = [x.cpu().pin_memory() for x in xs]
        h_loss = torch.empty(1, dtype=torch.float32).pin_memory()

        def e2e(i):
            o = tr.step(hx[i % 4])
            h_loss.copy_(o["loss"].reshape(-1)[:1], non_blocking=True)
        for i in range(2):
            e2e(i)
        e_ms = c.timed(e2e, steps) / steps
        loss = float(tr.out["loss"].reshape(-1)[0])
        for vqm in (model.vq_model_l, model.vq_model_h):
            vqm._codebook.check_peer_errors()
        res = {"batch_per_gpu": batch, "ms_per_step": ms, "traj_per_sec": world * batch / (ms * 1e-3),
               "e2e_ms_per_step": e_ms, "e2e_traj_per_sec": world * batch / (e_ms * 1e-3),
               "h2d_bytes_per_step": batch * 4 * 200 * 4, "d2h_bytes_per_step": 4, "loss_after": loss}
        if key == "weak":
            res["vq_share"] = vq_us_per_step * 1e-3 / ms
            res["vq_us_per_step"] = vq_us_per_step
        del tr, model
        torch.cuda.empty_cache()
        return res

    out["weak"] = one(B_TRAJ, "weak")
    if world > 1:
        out["strong"] = one(B_TRAJ // world, "strong")
    if rank == 0 and world == 1 and not args.no_cpu:
        out["cpu_reference"] = run_cpu_stage1()
    return out


# ------------------------------------------------------------------------------------------------ configs[2] sweep

def sweep_all(tvq, dev, hbm_gbs, bf16_tf, quick, note):
    ks = (512, 1024, 2048, 4096, 8192, 16384)
    ds = (64, 128, 256)
    big = {(512, 64), (16384, 256)}                    # + (32, 128): also measured at 2^24
    pts = [(1 << 22, 32, 128), (1 << 24, 32, 128)]
    for k in ks:
        for d in ds:
            ns = [1 << 20] if quick else [1 << 20, 1 << 22]
            if (k, d) in big and not quick:
                ns.append(1 << 24)
            pts += [(n, k, d) for n in ns]
    out = []
    cpu_cache = {}
    for (n, k, d) in pts:
        try:
            if (k, d) not in cpu_cache:
                cpu_cache[(k, d)] = cpu_per_latent(k, d)
            r = sweep_point(tvq, dev, n, k, d, hbm_gbs, bf16_tf)
            r["cpu_per_latent"] = cpu_cache[(k, d)]
            out.append(r)
            note(f"sweep {n}x{k}x{d}: " + ", ".join(f"{m['mode']} {m['ms']:.3f} ms frac {m['frac']:.3f}" for m in r["modes"]))
        except Exception as exc:
            out.append({"n": n, "k": k, "d": d, "error": f"{type(exc).__name__}: {exc}"[:200]})
        torch.cuda.empty_cache()
    return out


def cpu_per_latent(k, d):
    """Reference formula on the host CPU (oracle/vq_oracle.py: the reference's vq.py:210-218 row-chunked) at a reduced N,
    reported per latent: N * K * 4 bytes of `dist` would not fit host RAM at the sweep sizes."""
    import vq_oracle as O
    n = int(max(2048, min(1 << 16, 1.5e10 // (2 * k * d))))
    g = torch.Generator().manual_seed(1)
    x, e = torch.randn(n, d, generator=g), torch.randn(k, d, generator=g)
    torch.set_num_threads(os.cpu_count() or 1)
    O.assign_chunked(x[:256], e)
    t0 = time.perf_counter()
    O.assign_chunked(x, e)
    dt = time.perf_counter() - t0
    return {"ns_per_latent": dt / n * 1e9, "latents_per_sec": n / dt, "n": n, "mode": "eval_assign", "cores": os.cpu_count() or 1,
            "kind": "port"}


def sweep_point(tvq, dev, n, k, d, hbm_gbs, bf16_tf):
    """Isolated quantise at a BASELINE configs[2] size, three modes:
       eval_assign   indices only                       4d + 8 bytes / latent
       train_forward assign + q_st + loss + EMA update  8d + 8
       fwd_bwd       train_forward + backward           8d + 8 + 12d + 8
    each 2 k d FLOP / latent; bound = the slower roofline for the shape, frac = that roofline's time / ours."""
    lib = tvq._lib.load()
    g = torch.Generator(device=dev).manual_seed(1)
    nsets = 2 if n * d * 4 >= (1 << 28) else 3                  # distinct inputs: > 2 x L2 between reuses
    xs = [torch.randn(n, d, device=dev, generator=g) for _ in range(nsets)]
    e0 = torch.randn(k, d, device=dev, generator=g)
    gq = torch.randn(n, d, device=dev, generator=g) if n * d * 4 < (1 << 33) else xs[0]
    ws = tvq.Workspace(k, d, dev)
    idx = torch.empty(n, dtype=torch.int64, device=dev)
    q = torch.empty(n, d, device=dev)
    gx = torch.empty(n, d, device=dev)
    scal = torch.empty(8, device=dev)
    emb, csz, eavg, prev = e0.clone(), torch.zeros(k, device=dev), e0.clone(), torch.empty_like(e0)
    one = torch.ones(1, device=dev)

    def st():
        return torch.cuda.current_stream().cuda_stream

    def eval_assign(i):
        assert lib.tvq_forward(xs[i % nsets].data_ptr(), e0.data_ptr(), n, k, d, 0, 1.0, idx.data_ptr(), None, ws.stats.data_ptr(),
                               scal.data_ptr(), ws.buf.data_ptr(), ws.nbytes, st()) == 0

    def train_forward(i):
        emb.copy_(e0)        # keep the codebook fixed across repetitions (k*d*4 bytes: noise next to n*d*8)
        assert lib.tvq_train_step(xs[i % nsets].data_ptr(), emb.data_ptr(), csz.data_ptr(), eavg.data_ptr(), prev.data_ptr(), n, k, d,
                                  1.0, 0.8, 1e-5, idx.data_ptr(), q.data_ptr(), scal.data_ptr(), None, None, ws.buf.data_ptr(),
                                  ws.nbytes, st()) == 0

    def fwd_bwd(i):
        train_forward(i)
        assert lib.tvq_backward(gq.data_ptr(), None, one.data_ptr(), xs[i % nsets].data_ptr(), idx.data_ptr(), prev.data_ptr(), n, k, d,
                                1.0, gx.data_ptr(), st()) == 0

    t_tc = 2.0 * n * k * d / (bf16_tf * 1e12)
    modes = []
    for name, fn, bytes_per in (("eval_assign", eval_assign, 4 * d + 8), ("train_forward", train_forward, 8 * d + 8),
                                ("fwd_bwd", fwd_bwd, 20 * d + 16)):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        reps = 5 if n >= (1 << 24) else 10
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(reps):
            fn(i)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        t_hbm = n * bytes_per / (hbm_gbs * 1e9)
        modes.append({"mode": name, "ms": ms, "latents_per_sec": n / (ms * 1e-3), "hbm_gbs": n * bytes_per / (ms * 1e-3) / 1e9,
                      "tflops": 2.0 * n * k * d / (ms * 1e-3) / 1e12, "bound": "hbm" if t_hbm >= t_tc else "tensor",
                      "frac": max(t_hbm, t_tc) / (ms * 1e-3)})
    tf = modes[1]
    return {"n": n, "k": k, "d": d, "modes": modes, "mode": "train_forward", "ms": tf["ms"], "latents_per_sec": tf["latents_per_sec"],
            "bound": tf["bound"], "frac": tf["frac"],
            "path": "tcgen05 tf32, resident codebook" if k <= 32 and d <= 128 else "tcgen05 bf16 nomination, streamed codebook",
            "peak": "measured copy GB/s / measured cuBLAS bf16 TFLOP/s (MEASURED_PEAKS.json)"}


# ------------------------------------------------------------------------------------------------ configs[4] generation

def generation_bench(c, tvq, quick):
    """BASELINE configs[4]: 10 000 trajectories, batch 32 (configs/config.yaml:89, utils/sample_utils.py:15-53), T = 10 LF + 1 HF
    MaskGIT iterations (config.yaml:44-46).  The prior transformer is out of scope (un-vendored x-transformers, SURVEY
    section 8c): its logits are synthetic tensors of the right shape, so what is timed is everything AFTER the transformer —
    the sampling / re-masking step of every iteration (maskgit_step kernel: models/maskgit.py:300-346), the token ->
    decoder-input gather (decode_tokens: :465-470) and both decoders (harness, eval mode) -> x = x_l + x_h.  Beside it the
    same per-iteration step as the reference's eager torch ops (oracle/maskgit_oracle.py's op sequence on the GPU)."""
    import math
    import numpy as np
    dev = c.dev
    b, n_traj = 32, (1024 if quick else 10000)
    t_lf, t_hf, temp_lf, temp_hf = 10, 1, 10.0, 4.0
    torch.manual_seed(0); np.random.seed(0)
    model = tvq.Stage1(200, 4, tvq.stage1.default_config()).to(dev).eval()
    gen = torch.Generator(device=dev).manual_seed(5)
    logits_l = [torch.randn(b, TOK_LF, K_CODES, device=dev, generator=gen) for _ in range(t_lf)]
    logits_h = [torch.randn(b, TOK_HF, K_CODES, device=dev, generator=gen) for _ in range(t_hf)]

    def mask_len(t, T, n):          # cosine schedule (models/maskgit.py:95-103, :332-338)
        ratio = (t + 1) / T
        return int(max(0, min(n - 1, math.floor(n * math.cos(ratio * math.pi / 2))))) if t + 1 < T else 0

    def one_batch(i):
        s_l = torch.full((b, TOK_LF), K_CODES, dtype=torch.int64, device=dev)
        for t in range(t_lf):
            s_l = tvq.maskgit_step(logits_l[t], s_l, K_CODES, mask_len(t, t_lf, TOK_LF), temp_lf * (1 - (t + 1) / t_lf))
        s_h = torch.full((b, TOK_HF), K_CODES, dtype=torch.int64, device=dev)
        for t in range(t_hf):
            s_h = tvq.maskgit_step(logits_h[t], s_h, K_CODES, mask_len(t, t_hf, TOK_HF), temp_hf * (1 - (t + 1) / t_hf))
        zq_l = tvq.decode_tokens(s_l, model.vq_model_l, 3, 6, strict=False)
        zq_h = tvq.decode_tokens(s_h, model.vq_model_h, 3, 25, strict=False)
        return model.decoder_l(zq_l) + model.decoder_h(zq_h)

    with torch.no_grad():
        x = one_batch(0)
        assert x.shape == (b, 4, 200) and bool(torch.isfinite(x).all())
        n_batches = (n_traj + b - 1) // b
        for i in range(3):
            one_batch(i)
        eager_ms = c.timed(one_batch, n_batches, local=True)
        g_ms = c.graph_timed(one_batch, 8, local=True, replays=max(1, n_batches // 8))

        # the sampling step alone: kernel vs the reference's eager op sequence (torch on the GPU)
        s0 = torch.full((b, TOK_LF), K_CODES, dtype=torch.int64, device=dev)

        def k_step(i):
            tvq.maskgit_step(logits_l[i % t_lf], s0, K_CODES, 9, 5.0)

        def torch_step(i):
            logits = logits_l[i % t_lf]
            probs = torch.softmax(logits, -1)
            sampled = torch.distributions.categorical.Categorical(logits=logits).sample()
            unknown = s0 == K_CODES
            sampled = torch.where(unknown, sampled, s0)
            sel = torch.gather(probs, -1, sampled.unsqueeze(-1)).squeeze(-1)
            sel = torch.where(unknown, sel, torch.full_like(sel, float("inf")))
            u = torch.zeros_like(sel).uniform_(0, 1)
            conf = torch.log(sel + 1e-5) + 5.0 * (-torch.log(-torch.log(u.clamp_min(1e-20)).clamp_min(1e-20)))
            out = []
            for row in range(b):                                  # the reference's Python loop over the batch (:259-265)
                ind = torch.topk(conf[row], k=9, largest=False).indices
                m = torch.zeros(TOK_LF, dtype=torch.bool, device=dev)
                m[ind] = True
                out.append(m)
            return torch.where(torch.stack(out), torch.full_like(sampled, K_CODES), sampled)
        for i in range(3):
            k_step(i); torch_step(i)
        k_us = c.timed(k_step, 200, local=True) / 200 * 1e3
        t_us = c.timed(torch_step, 20, local=True) / 20 * 1e3
    per_batch = (g_ms if g_ms is not None else eager_ms / n_batches)
    return {"trajectories": n_traj, "batch": b, "iterations_per_batch": t_lf + t_hf,
            "ms_per_batch_eager": eager_ms / n_batches, "ms_per_batch_graph": g_ms,
            "traj_per_sec": b / (per_batch * 1e-3), "seconds_for_all": per_batch * 1e-3 * n_batches,
            "sampling_step_us_kernel": k_us, "sampling_step_us_reference_ops_on_gpu": t_us,
            "what": "synthetic logits (prior transformer out of scope) -> 10 + 1 x maskgit_step -> decode_tokens x 2 -> decoders -> x"}


# ------------------------------------------------------------------------------------------------ configs[0]

def config0_bench(c, tvq, args):
    """BASELINE configs[0]: the B = 32 stage-1 forward (train mode) — here on the GPU through the harness, the unmodified
    reference on the host CPU beside it; plus the `codebook_dim: 64` variant (32 x 64 codebook behind nn.Linear projections,
    SURVEY section 8 Note 1)."""
    import numpy as np
    dev = c.dev
    out = {}
    for name, extra in (("codebook_32x128", {}), ("codebook_32x64_projected", {"codebook_dim": 64})):
        torch.manual_seed(0); np.random.seed(0)
        cfg = tvq.stage1.default_config()
        cfg["VQ-VAE"].update(extra)
        model = tvq.Stage1(200, 4, cfg).to(dev).train()
        x = torch.rand(32, 4, 200, device=dev, generator=torch.Generator(device=dev).manual_seed(0)) * 2 - 1

        def fwd(i):
            with torch.no_grad():
                model.forward((x, None))
        for i in range(3):
            fwd(i)
        ms = c.timed(fwd, 20, local=True) / 20
        g_ms = c.graph_timed(fwd, 4, local=True)
        out[name] = {"ms_forward_eager": ms, "ms_forward_graph": g_ms, "traj_per_sec": 32 / ((g_ms or ms) * 1e-3)}
    if not args.no_cpu:
        try:
            import yaml
            import ref_loader
            _, ref_s1, _, root = ref_loader.load()
            cfg = yaml.safe_load(open(os.path.join(root, "configs", "config.yaml")))
            torch.manual_seed(0); np.random.seed(0)
            torch.set_num_threads(os.cpu_count() or 1)
            m = ref_s1.Stage1(200, 4, cfg).train()
            xc = torch.rand(32, 4, 200, generator=torch.Generator().manual_seed(0)) * 2 - 1
            yc = torch.zeros(32, 1, dtype=torch.long)
            with torch.no_grad():
                m((xc, yc), batch_idx=1)
                t0 = time.perf_counter()
                for _ in range(5):
                    m((xc, yc), batch_idx=1)
                dt = (time.perf_counter() - t0) / 5
            out["cpu_reference"] = {"ms_forward": dt * 1e3, "traj_per_sec": 32 / dt, "cores": os.cpu_count() or 1, "kind": "reference"}
        except Exception as exc:
            out["cpu_reference"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
    return out


def frontend_bench(c, tvq, hbm_gbs):
    """the STFT LF/HF front end of the configs[1] batch (1024 x 4 x 200), one kernel."""
    dev = c.dev
    try:
        xt = [torch.rand(B_TRAJ, 4, 200, device=dev) * 2 - 1 for _ in range(8)]       # 8 x 56 MB of outputs > L2
        outs = [tvq.lf_hf_frontend(x, 4, want=("enc_in_l", "enc_in_h", "x_l", "x_h")) for x in xt]   # warm-up + allocation
        lib_f = tvq._lib.load()

        def fe(i):
            x, o = xt[i % 8], outs[i % 8]
            assert lib_f.tvq_frontend(x.data_ptr(), B_TRAJ, 4, 200, 4, None, o["enc_in_l"].data_ptr(), o["enc_in_h"].data_ptr(),
                                      o["x_l"].data_ptr(), o["x_h"].data_ptr(), torch.cuda.current_stream().cuda_stream) == 0
        fe_ms = c.graph_timed(fe, 40, local=True) or c.timed(fe, 40, local=True) / 40
        fe_bytes = B_TRAJ * 4 * (200 * 4 + 2 * (2 * 3 * 201 * 4) + 2 * 200 * 4)
        return {"what": "tvq_frontend: x (1024,4,200) -> enc_in_l, enc_in_h (1024,8,3,201), x_l, x_h (1024,4,200); "
                        "n_fft=4 (stage1.py:101-113, vq_vae.py:179-180)", "us_per_launch": fe_ms * 1e3,
                "algorithmic_bytes": fe_bytes, "hbm_gbs": fe_bytes / (fe_ms * 1e-3) / 1e9,
                "frac_of_hbm_peak": fe_bytes / (fe_ms * 1e-3) / 1e9 / hbm_gbs}
    except Exception as exc:
        return {"error": str(exc)}


def main():
    import faulthandler
    faulthandler.dump_traceback_later(600, exit=False, file=sys.stderr)     # a hang leaves a Python stack in the log
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--quick", action="store_true", help="smaller sweep (2^20 only) and 1024 generated trajectories")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-stage1", action="store_true")
    ap.add_argument("--defer", type=int, default=2, choices=[0, 1, 2],
                    help="N > 1: 0 = statistics exchange inside the forward kernel's last CTA (round 1); 1 / 2 = deferred to "
                         "tvq_ema_finalize_dp on a side stream (include/tvq.h: tvq_hint_defer_exchange)")
    ap.add_argument("--no-sm-split", action="store_true", help="LF and HF forward launches each take the whole GPU (round-1 behaviour)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        our_arm(args)


if __name__ == "__main__":
    main()
