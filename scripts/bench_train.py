#!/usr/bin/env python
"""cfg 4 (training step) timing of the hot path: fwd + bwd of the whole EncoderDecoder (emrt_b200/train.py::build_train_step;
--msda-only: the 4 encoder + 2 decoder MSDeformableAttention modules of round 1) at B = 16 tiles of 512x512 (rows = 86 016),
bf16 activations, fp32 master gradients in all-reduce buckets, then the bucketed gradient all-reduce (NCCL when launched under
torchrun; a no-op at world size 1).
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
    ap.add_argument("--msda-only", action="store_true", help="round-1 definition: the six MSDeformableAttention modules on token inputs")
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
    from emrt_b200.train import build_train_step
    B = args.batch
    fwd_bwd, buckets, what = build_train_step(dev, rank, B, whole_model=not args.msda_only)

    def step():
        fwd_bwd()
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
        print(json.dumps({"metric": "training step (fwd + bwd + gradient all-reduce), tiles/s", "what": what,
                          "value": world * B / (ms * 1e-3), "unit": "512x512 tiles/s", "n_gpus": world, "ms_per_step": ms,
                          "batch_per_gpu": B, "grad_bucket_bytes": buckets.nbytes, "cuda_graph": bool(args.graph and world == 1),
                          "gpu_launches_per_step": ops.launch_count() // args.steps}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
