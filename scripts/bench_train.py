#!/usr/bin/env python
"""cfg 4 (training step) timing of the hot path: fwd + bwd of the 4 encoder + 2 decoder MSDeformableAttention modules at
B = 16 tiles of 512x512 (rows = 86 016), bf16 activations, fp32 master gradients in all-reduce buckets, then the bucketed
gradient all-reduce (NCCL when launched under torchrun; a no-op at world size 1).
  python scripts/bench_train.py [--batch 16] [--steps 5]            (or under torchrun for N > 1)
Prints one JSON line on rank 0.  Not the round's headline bench (bench.py is); this is the measurement of §8 row e."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import emrt_b200  # noqa: E402
from emrt_b200 import ops, synthetic  # noqa: E402
from emrt_b200.train import GradientBuckets  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--graph", action="store_true",
                    help="capture the step (forward, backward, bucket zeroing; world size 1 only) in one CUDA graph and time "
                         "replays: the eager step is bound by ~270 host-side launches, not by the kernels")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    shapes = synthetic.level_shapes(512)
    Lv = sum(h * w for h, w in shapes)
    B, C, Nq = args.batch, 256, 110
    mods = []
    for i in range(6):
        m = emrt_b200.MSDeformableAttention(C, 8, 3, 6).to(dev)
        with torch.no_grad():
            for name, arr in synthetic.msda_state(1234 + i).items():
                mod, leaf = name.split(".")
                getattr(getattr(m, mod), leaf).copy_(torch.from_numpy(arr))
        mods.append(m)
    buckets = GradientBuckets([p for m in mods for p in m.parameters()])
    g = torch.Generator(device="cpu").manual_seed(rank)
    src = torch.randn((B, Lv, C), generator=g).bfloat16().to(dev)
    tgt = torch.randn((B, Nq, C), generator=g).bfloat16().to(dev)
    pos = torch.randn((1, Lv, C), generator=g).bfloat16().to(dev)
    qpos = torch.randn((1, Nq, C), generator=g).bfloat16().to(dev)
    d_mem = torch.randn((B, Lv, C), generator=g).bfloat16().to(dev)
    d_hs = torch.randn((B, Nq, C), generator=g).bfloat16().to(dev)
    ref_enc = emrt_b200.get_reference_points(shapes, device=dev)
    ref_dec = torch.rand((1, Nq, 1, 2), generator=g).expand(-1, -1, 3, -1).contiguous().to(dev)

    def step():
        buckets.zero()
        x = src.clone().requires_grad_(True)
        mem = x
        for m in mods[:4]:
            mem = m(mem + pos, ref_enc, mem, shapes)       # with_pos_embed: torch add here (autograd glue, not timed apart)
        t = tgt.clone().requires_grad_(True)
        hs = t
        for m in mods[4:]:
            hs = m(hs + qpos, ref_dec, mem, shapes)
        torch.autograd.backward([mem, hs], [d_mem, d_hs])
        buckets.all_reduce()

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    run = step
    if args.graph and world == 1:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                      # warm the capture stream's allocator pool
            step()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
        run = graph.replay
        run()
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ops.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        print(json.dumps({"metric": "MSDA hot path training step (fwd+bwd of 6 modules + gradient all-reduce), tiles/s",
                          "value": world * B / (ms * 1e-3), "unit": "512x512 tiles/s", "n_gpus": world, "ms_per_step": ms,
                          "batch_per_gpu": B, "grad_bucket_bytes": buckets.nbytes, "cuda_graph": bool(args.graph and world == 1),
                          "gpu_launches_per_step": ops.launch_count() // args.steps}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
