"""Sub-records bench.py measures AFTER its main timing (each a few seconds; `--no-extras` skips them).  They put the
other BASELINE.json configurations and the platform ceilings in front of the driver, on the same box and at the same N:

  h2d_ceiling     pure pinned-memory host->device copy rate at the run's N (every rank copying at once) — what bounds `e2e`
  train_step      cfg 4: forward + backward of the hot path at 16 tiles of 512x512 per GPU, then the bucketed NCCL
                  all-reduce of the fp32 gradient buckets (train.py:116-123,153); the all-reduce timed apart, bus GB/s
  scene6000       cfg 5: one 6000x6000 scene, 256 windows, split into N bands of window rows (neighbour rows recomputed,
                  no collective on the data path); labels compared with the single-GPU stitch
  tiles256        cfg 2: 64 tiles of 256x256 through EncoderDecoder + head tail, tiles sharded over the N GPUs
  sensitivity     the encoder gather's time as the learned part of the sampling offsets grows (sigma of
                  sampling_offsets.weight) and with the window-centre hint off

Everything is timed with CUDA events on the launching stream, max over ranks.
"""
from __future__ import annotations

import os
import time

import torch
import torch.distributed as dist

TILE = 512


def _max_over_ranks(ms, dev, world):
    if world == 1:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _sync(world):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def _timed(fn, steps, warmup, dev, world):
    for _ in range(warmup):
        fn()
    _sync(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    _sync(world)
    return _max_over_ranks(e0.elapsed_time(e1) / steps, dev, world)


# ----------------------------------------------------------------------------------------------------------------
def h2d_ceiling(host_flat, dev_flat, dev, world, reps=8):
    """GB/s per GPU of `reps` back-to-back copies of the step's whole (contiguous, pinned) input buffer with nothing else
    running, all ranks at once: the platform's ceiling for the e2e leg's host->device traffic at this N."""
    ms = _timed(lambda: dev_flat.copy_(host_flat, non_blocking=True), reps, 2, dev, world)
    nbytes = host_flat.numel() * host_flat.element_size()
    return nbytes / (ms * 1e-3) / 1e9, ms


# ----------------------------------------------------------------------------------------------------------------
def build_encoder_decoder(dev, nclass, seed=1234, train=False):
    import emrt_b200
    from emrt_b200 import synthetic
    m = emrt_b200.EncoderDecoder(hidden_dim=256, dim_feedforward=1024, backbone_num_channels=[512, 1024, 2048], dropout=0.1,
                                 activation="relu", num_feature_levels=3, nhead=8, num_encoder_layers=4, num_decoder_layers=2,
                                 num_encoder_points=6, num_decoder_points=6, nclass=nclass)
    st = synthetic.encoder_decoder_state(seed)
    with torch.no_grad():
        sd = m.state_dict()
        for k in sd:
            sd[k].copy_(torch.from_numpy(st[k]))
    m = m.to(dev)
    return m if train else m.requires_grad_(False)


def _features(B, tile, dev, gen):
    r = lambda shape, std: (torch.randn(shape, generator=gen, device=dev) * std).bfloat16()
    feats = [r((B, c, tile // s, tile // s), 0.5) for c, s in zip((512, 1024, 2048), (8, 16, 32))]
    return feats, r((B, 256, 110), 0.5)


def train_step(dev, rank, world, batch=16, steps=4, warmup=2):
    """cfg 4.  Returns the sub-record.  `what` says which modules carry gradients."""
    import emrt_b200
    from emrt_b200 import ops, synthetic
    from emrt_b200.train import GradientBuckets, build_train_step
    step, buckets, what = build_train_step(dev, rank, batch)
    ms_eager = _timed(lambda: (step(), buckets.all_reduce()), steps, warmup, dev, world)
    # The eager step is bound by ~700 host-side launches (388 of this library's kernels + torch's gradient accumulation),
    # not by the kernels: capture forward + backward + bucket zeroing in ONE CUDA graph (the all-reduce stays outside it,
    # on NCCL's own stream) and time replays.  Falls back to the eager step if capture is not possible.
    run, graphed = step, False
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
        graph.replay()
        torch.cuda.synchronize()
        run, graphed = graph.replay, True
    except Exception:
        torch.cuda.synchronize()
    ms = _timed(lambda: (run(), buckets.all_reduce()), steps, warmup, dev, world) if graphed else ms_eager
    ms_compute = _timed(run, steps, 1, dev, world)
    rec = {"what": what, "batch_per_gpu": batch, "ms": ms, "ms_eager": ms_eager, "ms_fwd_bwd": ms_compute, "cuda_graph": graphed,
           "tiles_per_s": world * batch / (ms * 1e-3), "bytes": buckets.nbytes, "buckets": len(buckets.buckets)}
    if world > 1:
        ar = _timed(buckets.all_reduce, 10, 3, dev, world)
        rec["allreduce_ms"] = ar
        rec["bus_gbs"] = 2.0 * (world - 1) / world * buckets.nbytes / (ar * 1e-3) / 1e9
        # the reference's whole model is ~56 M parameters = 224 MB of fp32 gradients in ~25 MB buckets: the
        # same bucketed in-place all-reduce at that size, so the NVLink figure is not hidden behind a latency-bound 6-53 MB
        big = [torch.zeros(int(25e6 / 4), dtype=torch.float32, device=dev) for _ in range(9)]

        def big_ar():
            works = [dist.all_reduce(b, op=dist.ReduceOp.AVG, async_op=True) for b in big]
            for w in works:
                w.wait()
        t = _timed(big_ar, 10, 3, dev, world)
        nb = sum(b.numel() * 4 for b in big)
        rec["allreduce_full_model"] = {"bytes": nb, "buckets": len(big), "ms": t,
                                       "bus_gbs": 2.0 * (world - 1) / world * nb / (t * 1e-3) / 1e9,
                                       "nvlink_reference_gbs": 725.0}
    else:
        rec["allreduce_ms"], rec["bus_gbs"] = 0.0, None
    return rec


# ----------------------------------------------------------------------------------------------------------------
def scene6000(dev, rank, world, nc=6, window_batch=64):
    """cfg 5: the windows of this rank's band (owned window rows + the neighbours' rows that cover its label rows) go through
    EncoderDecoder.forward (synthetic C3-C5 features, `window_batch` windows per call) and the fused stitch; the bands are
    then summed into one map (disjoint rows: a 36 MB all-reduce AFTER the timed region, only for the check) and compared
    with rank 0's stitch of all 256 windows."""
    import emrt_b200
    from emrt_b200 import ops, sharding
    from emrt_b200.infer import window_origins
    H = W = 6000
    crop, stride = 512, 384
    plan, _, _ = emrt_b200.plan_windows([(H, W)], (crop, crop), (stride, stride))
    rows = window_origins(H, crop, stride)
    own, y0, y1, halo = sharding.shard_scene_rows(rows, crop, rank, world)
    use = sorted(set(own) | set(halo))
    sel = [k for k, p in enumerate(plan) if rows.index(p[1]) in use]
    g = torch.Generator(device=dev).manual_seed(6000)                        # same on every rank: window k's logits are rank-free
    half_all = torch.randn((len(plan), nc, crop // 2, crop // 2), generator=g, device=dev, dtype=torch.float32).bfloat16()
    model = build_encoder_decoder(dev, nc)
    feats, psp = _features(min(window_batch, len(sel)), TILE, dev, g)
    ti = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)
    band_y0 = min(plan[k][1] for k in sel)
    band_h = max(plan[k][1] for k in sel) + crop - band_y0
    half = half_all[sel].contiguous()
    wy, wx, wi = ti([plan[k][1] - band_y0 for k in sel]), ti([plan[k][2] for k in sel]), ti([0] * len(sel))
    out = {}

    def run():
        with torch.no_grad():
            for s0 in range(0, len(sel), window_batch):
                n = min(window_batch, len(sel) - s0)
                model([f[:n] for f in feats], psp[:n])
            out["lab"] = ops.stitch_argmax_fused(half, wi, wy, wx, 1, band_h, W, label_dtype=torch.uint8)[0]
    ms = _timed(run, 2, 1, dev, world)
    full = torch.zeros((1, 1, H, W), dtype=torch.uint8, device=dev)
    full[0, 0, y0:y1] = out["lab"][0, 0, y0 - band_y0:y1 - band_y0]
    if world > 1:
        acc = full.to(torch.int32)
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
        full = acc.to(torch.uint8)
    equal = None
    if rank == 0:
        single = ops.stitch_argmax_fused(half_all, ti([0] * len(plan)), ti([p[1] for p in plan]), ti([p[2] for p in plan]),
                                         1, H, W, label_dtype=torch.uint8)[0]
        equal = bool(torch.equal(single, full))
    return {"ms": ms, "windows": len(plan), "windows_this_rank_incl_halo": len(sel), "bands": world,
            "mpix_per_s": H * W / 1e6 / (ms * 1e-3), "labels_equal_single_gpu": equal,
            "what": "EncoderDecoder.forward on every window of the band + fused stitch/argmax of the band; halo window rows recomputed"}


# ----------------------------------------------------------------------------------------------------------------
def tiles256(dev, rank, world, total=64, nc=6):
    """cfg 2: `total` 256x256 tiles, contiguous slices per rank (shard_range), EncoderDecoder.forward + head tail."""
    from emrt_b200 import ops, sharding
    b, e = sharding.shard_range(total, rank, world)
    n = e - b
    g = torch.Generator(device=dev).manual_seed(256 + rank)
    model = build_encoder_decoder(dev, nc)
    feats, psp = _features(n, 256, dev, g)
    half = torch.randn((n, nc, 128, 128), generator=g, device=dev).bfloat16()
    idx = torch.arange(n, dtype=torch.int32, device=dev)
    z = torch.zeros(n, dtype=torch.int32, device=dev)

    def run():
        with torch.no_grad():
            model(feats, psp)
            ops.stitch_argmax_fused(half, idx, z, z, n, 256, 256, label_dtype=torch.uint8)
    ms = _timed(run, 5, 2, dev, world)
    return {"ms": ms, "tiles": total, "tiles_this_rank": n, "images_per_s": total / (ms * 1e-3), "scaling": "strong"}


# ----------------------------------------------------------------------------------------------------------------
def gather_sensitivity(dev, B=72):
    """Encoder-gather time (CUDA events around the launch, 4 launches each) against the spread of the sampling offsets:
    sigma = std of sampling_offsets.weight (the data-dependent part; the bench's weights use 0.05), and with the
    window-centre hint off.  The window-staged kernel reads samples outside its staged window from global memory; the last
    row is the L1-path kernel a model with wide offsets would select (MSDeformableAttention.window_gather = False)."""
    import emrt_b200
    from emrt_b200 import ops, synthetic, _lib as L
    shapes = synthetic.level_shapes(TILE)
    Lv = sum(h * w for h, w in shapes)
    g = torch.Generator(device=dev).manual_seed(7)
    src = torch.randn((B, Lv, 256), generator=g, device=dev).bfloat16()
    pos = torch.randn((1, Lv, 256), generator=g, device=dev).bfloat16()
    ref = emrt_b200.get_reference_points(shapes, device=dev)
    rows = []
    for sigma, hint, window in ((0.05, True, True), (0.15, True, True), (0.3, True, True), (0.05, False, True), (0.3, True, False)):
        m = emrt_b200.MSDeformableAttention(256, 8, 3, 6).to(dev)
        m.window_gather = window                  # False: the L1-path kernel, whose time does not depend on the spread
        with torch.no_grad():
            for name, arr in synthetic.msda_state(1234, offset_std=sigma).items():
                mod, leaf = name.split(".")
                getattr(getattr(m, mod), leaf).copy_(torch.from_numpy(arr))
        m.requires_grad_(False)
        if not hint:
            os.environ["EMRT_WIN_NO_HINT"] = "1"
        try:
            with torch.no_grad():
                m(src, ref, src, shapes, query_pos=pos)
                torch.cuda.synchronize()
                ops.kernel_events = []
                for _ in range(4):
                    m(src, ref, src, shapes, query_pos=pos)
                torch.cuda.synchronize()
                ev, ops.kernel_events = ops.kernel_events, None
        finally:
            os.environ.pop("EMRT_WIN_NO_HINT", None)
        t = [e[0].elapsed_time(e[1]) for (name, dims, e) in ev if name == "msda_gather_fwd"]
        rows.append({"offset_weight_sigma": sigma, "hint": hint, "kernel": "window-staged" if window else "l1-path (window_gather=False)",
                     "gather_ms": sum(t) / len(t)})
    return rows
