#!/usr/bin/env python
"""bench.py — EMRT hot path on B200: images/s (512x512 windows) and MPix/s.

A "step" = one pass of the hot path over one batch of synthetic input:
  the reference's whole EncoderDecoder.forward(src_feats, src_psp) (transformer_encoder_decoder.py:416-473, built with
  EMRT's own constructor arguments, paddle_EMRT.py:241-249) on the ResNet-50 C3-C5 feature maps of W 512x512 windows —
  input_proj (1x1 conv + GroupNorm), position / level embedding, 4 encoder layers (3x3 conv branch + GroupNorm + GELU,
  MSDeformableAttention [value / offset / weight / output projections, softmax over levels x points, multiscale bilinear
  gather], LayerNorm, FFN 256-1024-256, LayerNorm), 2 decoder layers (110-token self-attention, MSDeformableAttention
  cross-attention, FFN) — then the head tail (x2 bilinear upsample + sliding-window overlap stitch + softmax + argmax)
  that turns the windows' half-resolution class logits into 1024x1024 LoveDA-shaped label maps (9 windows per image,
  window 512, stride 384).
Out of the step (out of scope): ResNet-50 backbone, spatial branch / PSP, EFP, conv heads.
  --tokens     the earlier step definition (token inputs: TransformerEncoder + 2 bare decoder MSDA calls + head tail)
  --msda-only  the round-1 starting definition (4 + 2 bare MSDA calls + head tail)

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (one rank per GPU under torchrun)
  python bench.py --impl reference ...                           the reference path's CPU restatement (oracle)
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TILE = 512
NC = 7                      # LoveDA
WINDOWS_PER_IMAGE = 9       # 1024x1024, window 512, stride 384 -> origins {0, 384, 512}^2
SCENE = 1024
METRIC = ("images/s (512x512 windows through the EMRT hot path: EncoderDecoder.forward on the C3-C5 features [input_proj, "
          "4 encoder layers with conv branch + MSDeformableAttention + LayerNorm + FFN, 2 decoder layers with self-attention + "
          "MSDeformableAttention + FFN] + upsample/stitch/argmax)")
FEAT_CH = (512, 1024, 2048)   # ResNet-50 C3..C5 (paddle_EMRT.py:191)
WORKLOAD_TAIL = {"full": "hot path only (EncoderDecoder.forward on C3-C5 features + head tail)",
                 "tokens": "hot path only (4-layer encoder + 2 decoder MSDA + head tail; token inputs)",
                 "msda": "4 + 2 bare MSDeformableAttention calls + head tail"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images", type=int, default=8, help="1024x1024 scenes per GPU per step (9 windows each)")
    ap.add_argument("--gemm", default="auto", choices=["auto", "simt", "tcgen05"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the sub-records measured after the main timing "
                    "(cfg-4 training step, cfg-5 scene, cfg-2 tiles, copy ceiling, gather sensitivity)")
    ap.add_argument("--cpu-sample-windows", type=int, default=8)
    ap.add_argument("--e2e-sub-batches", action="store_true", help="e2e: 4 sub-batches of whole scenes per step (shorter pipeline fill, "
                    "4x the launches: measured slower, 11.56 vs 11.28 ms per step at 20 steps) instead of one copy + one pass per step")
    ap.add_argument("--msda-only", action="store_true", help="round-1 starting definition: 4 + 2 bare MSDA calls + head tail")
    ap.add_argument("--tokens", action="store_true", help="earlier definition: token inputs, encoder + 2 bare decoder MSDA + head tail")
    a = ap.parse_args()
    a.mode = "msda" if a.msda_only else ("tokens" if a.tokens else "full")
    return a


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                clk = float(f[1]); mx = float(f[2])
            except ValueError:
                continue
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(clk)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:      # timed region shorter than the sampling period: use every sample we have
            for ts, line in self.lines:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[1])); mx = float(f[2])
                except Exception:
                    pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_traffic(B, live=True):
    """DRAM bytes per launch of the dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum) and where the figure
    comes from.  `live`: measured NOW, by running scripts/gather_once.py (the same kernel on the same geometry) under ncu
    in a subprocess after the timed region — nothing timed runs under the profiler; otherwise, or when ncu is unavailable,
    the committed capture profiles/traffic.json (keyed by windows per launch)."""
    if live:
        try:
            cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "--csv",
                   "-k", "regex:msda_gather_fwd_win", "-s", "1", "-c", "1", sys.executable,
                   os.path.join(ROOT, "scripts", "gather_once.py"), str(B)]
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
            tot, seen = 0.0, 0
            for line in r.stdout.splitlines():
                f = [x.strip('"') for x in line.split('","')]
                if len(f) > 3 and f[-3] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[f[-2]]
                    tot += float(f[-1].replace(",", "")) * scale
                    seen += 1
            if seen == 2:
                return tot, "ncu subprocess in this run (scripts/gather_once.py, same kernel and geometry)"
        except Exception:
            pass
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        d = json.load(open(p))
        e = d["msda_gather_fwd_win"].get(str(B))
        return (None, None) if e is None else (float(e["dram_bytes_read"]) + float(e["dram_bytes_write"]), "profiles/traffic.json (committed capture)")
    except Exception:
        return None, None


def bind_to_gpu_numa_node(local):
    """Pin this rank's CPU threads to the cores NVML reports as local to its GPU, BEFORE the pinned host buffers are
    allocated (first touch puts them on that NUMA node): with 8 ranks each streaming ~0.6 GB per step over PCIe, host
    buffers on the far socket halve the e2e rate.  Best effort: silently skipped when NVML / affinity is unavailable."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local).uuid)
        h = None
        for cand in (uuid, "GPU-" + uuid):
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID(cand.encode() if isinstance(cand, str) else cand)
                break
            except Exception:
                continue
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64 + 1)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        pick = cpus & allowed
        if pick and pick != allowed:
            os.sched_setaffinity(0, pick)
            return len(pick)
    except Exception:
        pass
    return 0


def window_tables(n_img):
    import emrt_b200
    plan, H, W = emrt_b200.plan_windows([(SCENE, SCENE)] * n_img, (TILE, TILE), (384, 384))
    assert len(plan) == n_img * WINDOWS_PER_IMAGE
    return plan, H, W


def gather_bytes(B, Lq, Lv, M, D, L, P, sv, sl):
    """Algorithmic bytes of one gather launch (BASELINE.md §3 / SURVEY.md §8d)."""
    return sv * B * Lv * M * D + sl * B * Lq * M * L * P * 3 + sv * B * Lq * M * D


# ------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference path's CPU restatement (the oracle), on the host cores
# ------------------------------------------------------------------------------------------------------------
def cpu_hot_path(n_windows, threads, repeats=1, warmup=0, mode="full"):
    """Times the oracle (torch-CPU fp32 restatement of the reference's Paddle op composition, checked against the
    reference's own sources in tests/test_reference_pin.py) on `n_windows` 512x512 windows of the same workload.
    Returns (images/s, seconds per pass)."""
    import numpy as np
    import torch
    import oracle as O
    from emrt_b200 import synthetic
    torch.set_num_threads(threads)
    if mode == "full":
        rng = np.random.Generator(np.random.PCG64(0))
        params = {k: torch.from_numpy(v) for k, v in synthetic.encoder_decoder_state(1234).items()}
        feats = [torch.from_numpy(O.rng_normal(rng, (n_windows, c, TILE // s, TILE // s), 0.5)) for c, s in zip(FEAT_CH, (8, 16, 32))]
        psp = torch.from_numpy(O.rng_normal(rng, (n_windows, 256, 110), 0.5))
        half = torch.from_numpy(O.rng_normal(rng, (n_windows, NC, TILE // 2, TILE // 2)))

        def one_pass():
            hs, mem, _ = O.encoder_decoder_forward(params, feats, psp, num_enc=4, num_dec=2)
            full = O.upsample2x(half)                       # UpHead tail
            return O.ss_inference_tail(full, (TILE, TILE)), hs
        with torch.no_grad():
            for _ in range(warmup):
                one_pass()
            t0 = time.perf_counter()
            for _ in range(repeats):
                one_pass()
            dt = (time.perf_counter() - t0) / repeats
        return n_windows / dt, dt
    shapes = [(TILE // 8,) * 2, (TILE // 16,) * 2, (TILE // 32,) * 2]
    _, Lv = O.level_tables(shapes)
    rng = np.random.Generator(np.random.PCG64(0))
    enc = {}
    for i in range(4):
        for k, v in synthetic.encoder_layer_state(1234 + 10 * i).items():
            enc[f"encoder.layers.{i}.{k}"] = torch.from_numpy(v)
    dec = [{k: torch.from_numpy(v) for k, v in O.make_msda_params(1234 + i).items()} for i in range(4, 6)]
    src = torch.from_numpy(O.rng_normal(rng, (n_windows, Lv, 256)))
    pos = torch.from_numpy(O.rng_normal(rng, (1, Lv, 256)))
    tgt = torch.from_numpy(O.rng_normal(rng, (n_windows, 110, 256)))
    qpos = torch.from_numpy(O.rng_normal(rng, (1, 110, 256)))
    half = torch.from_numpy(O.rng_normal(rng, (n_windows, NC, TILE // 2, TILE // 2)))
    ref_enc = O.encoder_reference_points(shapes, n_windows)
    ref_dec = torch.rand(n_windows, 110, 1, 2).expand(-1, -1, 3, -1).contiguous()
    mask = torch.ones(n_windows, Lv)

    def one_pass():
        x = src
        for i in range(4):
            x = O.encoder_layer_forward(enc, f"encoder.layers.{i}.", x, ref_enc, shapes, mask, pos)
        t = tgt
        for i in range(2):
            t = O.msda_forward(dec[i], t + qpos, ref_dec, x, shapes, mask)
        full = O.upsample2x(half)                       # UpHead tail
        return O.ss_inference_tail(full, (TILE, TILE)), t   # softmax + argmax (stitching is index work only)

    with torch.no_grad():
        for _ in range(warmup):
            one_pass()
        t0 = time.perf_counter()
        for _ in range(repeats):
            one_pass()
        dt = (time.perf_counter() - t0) / repeats
    return n_windows / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    nwin = max(1, args.cpu_sample_windows)
    ips, dt = cpu_hot_path(nwin, cores, repeats=max(1, min(args.steps, 3)), warmup=1 if args.warmup > 0 else 0, mode=args.mode)
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus,
        "steps": max(1, min(args.steps, 3)), "warmup": 1 if args.warmup > 0 else 0, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "mpix_per_s": ips * TILE * TILE / 1e6,
        "config": {"workload": f"cfg3-shaped: {TILE}x{TILE} windows of 1024x1024 LoveDA scenes, {NC} classes, "
                               "window 512 stride 384; " + WORKLOAD_TAIL[args.mode],
                   "sample": f"each step = {nwin} of the {args.images * WINDOWS_PER_IMAGE} windows of our arm's step; steps clamped to "
                             f"{max(1, min(args.steps, 3))} and warm-up to {1 if args.warmup > 0 else 0} (requested {args.steps} / "
                             f"{args.warmup}) so the CPU arm ends within minutes",
                   "note": "reference = CPU restatement of the reference's Paddle op composition (oracle/, torch-CPU "
                           "fp32, verified equal to the reference's own sources in tests/test_reference_pin.py); "
                           "PaddlePaddle itself cannot be installed in this image"},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{nwin} window(s) of the {args.images * WINDOWS_PER_IMAGE}-window step, all host threads"},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import emrt_b200
    from emrt_b200 import ops, _lib as L
    from emrt_b200.hotpath import HotPath

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ops.device_check()
    numa_cpus = bind_to_gpu_numa_node(local) if world > 1 else 0

    impl = {"auto": L.IMPL_AUTO, "simt": L.IMPL_SIMT, "tcgen05": L.IMPL_TCGEN05}[args.gemm]
    hp = HotPath(dev, TILE, NC, gemm_impl=impl, mode=args.mode)

    def set_conv_overlap(k):
        """TransformerEncoderLayer.overlap_conv of every layer: k > 0 runs the 3x3 conv on k SMs of a side stream beside the
        sampling gather (the library's default); 0 keeps the layer on one stream, in order."""
        root = hp.model if hp.model is not None else hp.encoder
        layers = [m for m in root.modules() if hasattr(m, "overlap_conv")] if root is not None else []
        prev = layers[0].overlap_conv if layers else 0
        for m in layers:
            m.overlap_conv = k
        return prev
    # The timed regions run the layers IN ORDER on one stream: the per-kernel CUDA-event durations behind `roofline` and
    # `dense_kernels` are then those of each kernel alone (beside the conv, the gather's launch lasts 0.99 ms instead of 0.73 —
    # it is sharing the SMs — although the pair finishes sooner).  The overlapped form is timed after them: `conv_beside_gather`.
    lib_overlap = set_conv_overlap(0)
    full = args.mode == "full"
    n_img = args.images
    B = n_img * WINDOWS_PER_IMAGE
    plan, H, W = window_tables(n_img)
    Lv, C, Nq = hp.Lv, hp.C, hp.num_queries
    ti = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)
    win = dict(win_img=ti([p[0] for p in plan]), win_y0=ti([p[1] for p in plan]), win_x0=ti([p[2] for p in plan]),
               n_img=n_img, H=H, W=W)
    g = torch.Generator(device="cpu").manual_seed(1000 + rank)
    pos = torch.randn((1, Lv, C), generator=g).bfloat16().to(dev)
    qpos = torch.randn((1, Nq, C), generator=g).bfloat16().to(dev)
    mask = torch.ones((B, Lv), dtype=torch.float32, device=dev)

    # host (pinned) inputs for the e2e path and N_SETS resident device copies for the kernel-only path; rotating
    # over N_SETS input sets (> 126 MB L2 in total) keeps every timed step's inputs cold.
    feat_elems = sum(c * (TILE // s) ** 2 for c, s in zip(FEAT_CH, (8, 16, 32)))
    set_bytes = 2 * ((B * feat_elems + B * C * Nq if full else B * Lv * C + B * Nq * C) + B * NC * (TILE // 2) ** 2)
    N_SETS = 4 if set_bytes < 150e6 else 2          # either way the rotating sets exceed the 126 MB L2

    def as_batch(d):
        """flat dict of tensors (what travels host -> device) -> the step's argument dict"""
        b = dict(d)
        if full:
            b["feats"] = [b.pop("c3"), b.pop("c4"), b.pop("c5")]
        b.update(pos=pos, qpos=qpos, mask=mask, **win)
        return b
    # every input set is ONE contiguous pinned buffer (one host->device copy per step); the tensors are views into it
    if full:        # ResNet-50 C3..C5 feature maps [B, 512|1024|2048, T/8|T/16|T/32, .] and the PSP tokens [B, 256, 110]
        layout = [(name, (B, c, TILE // st, TILE // st), 0.5) for name, c, st in zip(("c3", "c4", "c5"), FEAT_CH, (8, 16, 32))]
        layout.append(("psp", (B, C, Nq), 0.5))
    else:
        layout = [("src", (B, Lv, C), 1.0), ("tgt", (B, Nq, C), 1.0)]
    layout.append(("half_logits", (B, NC, TILE // 2, TILE // 2), 1.0))

    def views(flat):
        out, off = {}, 0
        for name, shape, _ in layout:
            n = int(np.prod(shape))
            out[name] = flat[off:off + n].view(shape)
            off += (n + 127) // 128 * 128                    # 256-byte aligned starts (TMA / 16-byte vector loads)
        return out
    flat_elems = sum((int(np.prod(sh)) + 127) // 128 * 128 for _, sh, _ in layout)
    host_flats, host_sets, dev_sets = [], [], []
    for s in range(N_SETS):
        hf = torch.empty(flat_elems, dtype=torch.bfloat16).pin_memory()
        hv = views(hf)
        for name, shape, std in layout:
            hv[name].copy_((torch.randn(shape, generator=g) * std).bfloat16())
        host_flats.append(hf)
        host_sets.append(hv)
        dev_sets.append(as_batch(views(hf.to(dev))))
    in_bytes = flat_elems * 2
    labels_dev = torch.empty((n_img, 1, H, W), dtype=torch.uint8, device=dev)
    labels_host = torch.empty((n_img, 1, H, W), dtype=torch.uint8).pin_memory()
    hs_host = torch.empty((B, Nq, C), dtype=torch.bfloat16).pin_memory()
    out_bytes = labels_host.numel() + hs_host.numel() * 2

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- kernel-only: inputs resident in HBM -------------------------------------------------------------
    with torch.no_grad():
        for i in range(args.warmup):
            hp.step(dev_sets[i % N_SETS], labels_dev)
        sync_all()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
            time.sleep(0.25)
        ops.prewarm_events(5 * 8 * args.steps)         # the events the fused MSDA calls of the timed region will record in place
        ops.kernel_events = []
        ops.reset_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        t_wall0 = time.time()
        e0.record()
        for i in range(args.steps):
            hp.step(dev_sets[i % N_SETS], labels_dev)
        e1.record()
        sync_all()
        t_wall1 = time.time()
        launches = ops.launch_count()
        events, ops.kernel_events = ops.kernel_events, None
        clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
        ms_total = max_over_ranks(e0.elapsed_time(e1))

        # ---- e2e: host buffers in, host labels out, every step -------------------------------------------
        # A step's input set is ONE contiguous pinned buffer -> one cudaMemcpyAsync into one of three device slots: the copies of
        # steps i+1, i+2 overlap the kernels of step i and the D2H of step i's labels / decoder states overlaps step i+1.  Every
        # step's copies are inside the timed region, which starts and ends with nothing in flight: K steps cost K copies + one
        # step of kernels (the drain) — 11.28 ms per step at K = 20 with a 10.8 ms copy.  --e2e-sub-batches sends the set as
        # four sub-batches of whole scenes instead (a quarter of the fill / drain, four times the launches: measured slower).
        cur = torch.cuda.current_stream()
        h2d, d2h = torch.cuda.Stream(), torch.cuda.Stream()
        NSLOT = 3
        NSUB = 4 if (full and n_img % 4 == 0 and args.e2e_sub_batches) else 1
        n_img_s, B_s = n_img // NSUB, B // NSUB
        plan_s, H_s, W_s = window_tables(n_img_s)
        win_s = dict(win_img=ti([p[0] for p in plan_s]), win_y0=ti([p[1] for p in plan_s]), win_x0=ti([p[2] for p in plan_s]),
                     n_img=n_img_s, H=H_s, W=W_s)
        layout_s = [(name, (B_s,) + tuple(shape[1:]), std) for name, shape, std in layout]
        sub_elems = sum((int(np.prod(sh)) + 127) // 128 * 128 for _, sh, _ in layout_s)

        def views_s(flat):
            out, off = {}, 0
            for name, shape, _ in layout_s:
                n = int(np.prod(shape))
                out[name] = flat[off:off + n].view(shape)
                off += (n + 127) // 128 * 128
            return out

        def as_batch_s(d):
            b = dict(d)
            if full:
                b["feats"] = [b.pop("c3"), b.pop("c4"), b.pop("c5")]
            b.update(pos=pos, qpos=qpos, mask=mask[:B_s], **win_s)
            return b
        # the same input sets, re-laid per sub-batch (windows [j B_s, (j + 1) B_s) of every tensor) in pinned memory
        host_subs = []
        for sset in range(N_SETS):
            subs = []
            for j in range(NSUB):
                hf = torch.empty(sub_elems, dtype=torch.bfloat16).pin_memory()
                hv = views_s(hf)
                for name, _, _ in layout_s:
                    hv[name].copy_(host_sets[sset][name][j * B_s:(j + 1) * B_s])
                subs.append(hf)
            host_subs.append(subs)
        slot_flats = [torch.empty(sub_elems, dtype=torch.bfloat16, device=dev) for _ in range(NSLOT)]
        slots = [views_s(f) for f in slot_flats]
        lab_slots = [torch.empty((n_img_s, 1, H_s, W_s), dtype=torch.uint8, device=dev) for _ in range(NSLOT)]
        ready = [torch.cuda.Event() for _ in range(NSLOT)]
        free = [torch.cuda.Event() for _ in range(NSLOT)]
        done = [torch.cuda.Event() for _ in range(NSLOT)]
        drained = [torch.cuda.Event() for _ in range(NSLOT)]
        hs_keep = [None] * NSLOT
        for ev in free + drained:
            ev.record(cur)
        sub_counter = [0]

        def e2e_step(i):
            for j in range(NSUB):
                g_ = sub_counter[0]
                sub_counter[0] += 1
                sl = g_ % NSLOT
                with torch.cuda.stream(h2d):
                    h2d.wait_event(free[sl])               # the kernels of sub-batch g - NSLOT are done with this slot
                    slot_flats[sl].copy_(host_subs[i % N_SETS][j], non_blocking=True)     # ONE copy: the whole sub-batch
                    ready[sl].record(h2d)
                cur.wait_event(ready[sl])
                cur.wait_event(drained[sl])                # sub-batch g - NSLOT's labels have left lab_slots[sl]
                lab, hs = hp.step(as_batch_s(slots[sl]), lab_slots[sl])
                hs_keep[sl] = hs
                free[sl].record(cur)
                done[sl].record(cur)
                with torch.cuda.stream(d2h):
                    d2h.wait_event(done[sl])
                    labels_host[j * n_img_s:(j + 1) * n_img_s].copy_(lab, non_blocking=True)
                    hs_host[j * B_s:(j + 1) * B_s].copy_(hs, non_blocking=True)
                    hs.record_stream(d2h)
                    drained[sl].record(d2h)

        def e2e_sync():
            h2d.synchronize(); d2h.synchronize()
        for i in range(min(args.warmup, 3)):
            e2e_step(i)
        e2e_sync()
        sync_all()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for i in range(args.steps):
            e2e_step(i)
        cur.wait_stream(d2h)                           # the last labels are on the host when f1 fires
        f1.record()
        e2e_sync()
        sync_all()
        ms_e2e = max_over_ranks(f0.elapsed_time(f1))

        # pure-copy ceiling of the platform at this N: the same buffer, the same copy, nothing else running
        import bench_extras as X
        ceil_flat = torch.empty(flat_elems, dtype=torch.bfloat16, device=dev)
        ceil_gbs, ceil_ms = X.h2d_ceiling(host_flats[0], ceil_flat, dev, world)
        del ceil_flat

    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)
    e2e_value = world * B / (ms_e2e / args.steps * 1e-3)
    copy_bound = world * B / (ceil_ms * 1e-3)                  # images/s if the step were nothing but its H2D copy
    e2e_rec = {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
               "ms_per_step": ms_e2e / args.steps, "h2d_gbs_per_gpu": in_bytes / (ms_e2e / args.steps * 1e-3) / 1e9,
               "h2d_ceiling_gbs": ceil_gbs, "h2d_ceiling_ms": ceil_ms,
               "h2d_ceiling_what": f"{in_bytes / 1e6:.0f} MB pinned -> device, one cudaMemcpyAsync, all {world} rank(s) at once, "
                                   "nothing else running (max over ranks)",
               "bound": "copy" if copy_bound < value else "compute",
               "frac_of_bound": e2e_value / min(value, copy_bound),
               "copies": f"{NSUB} sub-batch(es) of whole scenes per step, each one contiguous pinned buffer -> one copy, 3 device slots, "
                         "H2D / kernels / D2H on 3 streams"}

    # ---- the same steps with the conv beside the gather (two streams): what the library does by default ----
    overlap_rec = None
    if lib_overlap > 0 and args.mode == "full" and not args.no_extras:
        set_conv_overlap(lib_overlap)
        with torch.no_grad():
            for i in range(args.warmup):
                hp.step(dev_sets[i % N_SETS], labels_dev)
            sync_all()
            o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            o0.record()
            for i in range(args.steps):
                hp.step(dev_sets[i % N_SETS], labels_dev)
            o1.record()
            sync_all()
        ms_o = max_over_ranks(o0.elapsed_time(o1)) / args.steps
        overlap_rec = {"ms_per_step": ms_o, "value": world * B / (ms_o * 1e-3), "unit": "images/s", "conv_sms": lib_overlap,
                       "what": "the same K steps, inputs resident, with each encoder layer's 3x3 conv on a side stream beside its "
                               "sampling gather (TransformerEncoderLayer.overlap_conv, the library default; bit-equal outputs)"}
        set_conv_overlap(0)

    # ---- the other configurations and splits, short, after the main timing --------------------------------
    extras = {}
    if not args.no_extras and args.mode == "full":
        del dev_sets, slot_flats, slots, host_subs
        torch.cuda.empty_cache()
        for name, fn in (("train_step", lambda: X.train_step(dev, rank, world)),
                         ("scene6000", lambda: X.scene6000(dev, rank, world)),
                         ("tiles256", lambda: X.tiles256(dev, rank, world))):
            try:
                extras[name] = fn()
            except Exception as exc:                              # a sub-record must never take the headline down
                extras[name] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
                if world > 1:
                    raise
            torch.cuda.empty_cache()
        if world == 1:
            try:
                extras["sensitivity"] = X.gather_sensitivity(dev, B)
            except Exception as exc:
                extras["sensitivity"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    # ---- roofline of the dominant kernel (the encoder gather) --------------------------------------------
    enc = [(dims, ev[0].elapsed_time(ev[1])) for (name, dims, ev) in events if name == "msda_gather_fwd" and dims[1] == Lv]
    # the dense contractions of the step against both of their bounds (north_star: tensor-pipe fraction of the projections)
    peaks_d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    tf_peak = float(peaks_d.get("bf16_tflops_sustained", 1407.0))      # kernels timed inside a long step: sustained figure
    tf_burst = float(peaks_d.get("bf16_tflops", 1681.0))
    dense = {}
    for (name, dims, ev) in events:
        if name == "linear" and dims[0] == B * Lv:
            rows_, K_, N_, sx, sy = dims
            key = f"linear rows={rows_} K={K_} N={N_}"
            flops, byts = 2.0 * rows_ * K_ * N_, rows_ * (K_ * sx + N_ * sy) + K_ * N_ * sx
        elif name == "linear_ln" and dims[0] == B * Lv:
            rows_, K_, N_, sx, sy = dims
            key = f"linear + residual + LayerNorm rows={rows_} K={K_} N={N_} (output_proj + norm1)"
            flops, byts = 2.0 * rows_ * K_ * N_, rows_ * (K_ * sx + 2 * N_ * sy) + K_ * N_ * sx      # x and the residual in, y out
        elif name == "ffn_fused" and dims[0] == B * Lv:
            rows_, C_, F_, sx = dims
            key = f"ffn_fused rows={rows_} d_model={C_} d_ff={F_} (linear1 + ReLU + linear2 + residual + norm2 + conv branch)"
            # x, conv, skip in; y out; both weight matrices once
            flops, byts = 4.0 * rows_ * C_ * F_, 4.0 * rows_ * C_ * sx + 2.0 * C_ * F_ * sx
        elif name == "conv3x3" and dims[0] == B * Lv:
            rows_, C_, sx = dims
            key = f"conv3x3 rows={rows_} C={C_}"
            flops, byts = 2.0 * rows_ * 9 * C_ * C_, 2.0 * rows_ * C_ * sx + 3 * 9 * C_ * C_ * sx
        else:
            continue
        d = dense.setdefault(key, {"ms": 0.0, "n": 0, "flops": flops, "bytes": byts})
        d["ms"] += ev[0].elapsed_time(ev[1]); d["n"] += 1
    dense_rows = []
    for key, d in dense.items():
        t = d["ms"] / d["n"] * 1e-3
        dense_rows.append({"kernel": key, "launches_timed": d["n"], "avg_launch_ms": t * 1e3,
                           "tflops": d["flops"] / t / 1e12, "frac_of_bf16_peak": d["flops"] / t / 1e12 / tf_peak,
                           "frac_of_bf16_burst_peak": d["flops"] / t / 1e12 / tf_burst,
                           "gbs": d["bytes"] / t / 1e9, "frac_of_hbm_peak": d["bytes"] / t / 1e9 / peaks()[0]})
    hbm_peak, peak_src = peaks()
    roof = None
    if enc:
        avg_ms = sum(t for _, t in enc) / len(enc)
        gb = gather_bytes(*enc[0][0]) / 1e9
        ach = gb / (avg_ms * 1e-3)
        roof = {"kernel": "msda_gather_fwd (encoder call, Lq=Lv=%d, B=%d)" % (Lv, B), "bound": "hbm",
                "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None,
                "algorithmic_bytes": gb * 1e9, "avg_launch_ms": avg_ms, "launches_timed": len(enc),
                "share_of_step": avg_ms * len(enc) / args.steps / ms_step, "peak_source": peak_src}
        # what actually bounds it: every (query, head) reads L*P*4 corners x D x 2 B from shared memory
        b_, lq_, _, m_, d_, l_, p_, sv_, _ = enc[0][0]
        onchip = float(b_) * lq_ * m_ * l_ * p_ * 4 * d_ * sv_
        sm_clk = (clocks or {}).get("sm_mhz") or 1965.0
        pipe = torch.cuda.get_device_properties(dev).multi_processor_count * 128.0 * sm_clk * 1e6
        roof["on_chip"] = {"what": "shared-memory data pipe (128 B/clk/SM): corner reads of the gather, a floor for any formulation "
                                   "that reads each bilinear corner from on-chip memory",
                           "bytes": onchip, "achieved_gbs": onchip / (avg_ms * 1e-3) / 1e9, "peak_gbs": pipe / 1e9,
                           "frac": onchip / (avg_ms * 1e-3) / pipe}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "mpix_per_s": value * TILE * TILE / 1e6,
        "config": {"workload": f"cfg3-shaped: {B} {TILE}x{TILE} windows per GPU per step = {n_img} LoveDA 1024x1024 "
                               f"scenes (window 512, stride 384), {NC} classes; " + WORKLOAD_TAIL[args.mode],
                   "windows_per_gpu": B, "tokens_per_window": Lv, "gemm": args.gemm,
                   "l2": f"inputs rotate over {N_SETS} resident sets ({N_SETS * in_bytes / 1e6:.0f} MB > 126 MB L2); "
                         "each step also streams > 1 GB of intermediates",
                   "parallelism": f"windows sharded over {world} GPU(s), no collective",
                   "host_affinity": f"rank threads bound to {numa_cpus} GPU-local cores" if numa_cpus else "default"},
        "e2e": e2e_rec,
        "conv_beside_gather": overlap_rec,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "dense_kernels": {"peak_tflops": tf_peak, "burst_peak_tflops": tf_burst,
                          "peak_source": "MEASURED_PEAKS.json: bf16_tflops_sustained (cuBLAS back to back for 4 s — the figure for kernels "
                                         "timed inside a long step; a tile kernel can exceed it) and bf16_tflops (burst)",
                          "rows": dense_rows},
    }
    line.update({k: v for k, v in extras.items() if k != "sensitivity"})
    if roof is not None:
        if "sensitivity" in extras:
            roof["sensitivity"] = extras["sensitivity"]
        roof["traffic"], roof["traffic_source"] = measured_traffic(B, live=(world == 1 and not args.no_extras))
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        ips, dt = cpu_hot_path(max(1, args.cpu_sample_windows), cores, repeats=2, warmup=1, mode=args.mode)
        line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                                "sample": f"{max(1, args.cpu_sample_windows)} window(s) of the {B}-window step, "
                                          f"torch-CPU fp32 oracle, {cores} threads, {dt:.2f} s per pass"}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
