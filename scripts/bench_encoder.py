#!/usr/bin/env python
"""Timing of the full TransformerEncoder (4 layers: conv branch + MSDA + LN + FFN) on B windows of
512x512 (Lv = 5376), bf16.  python scripts/bench_encoder.py [--windows 18] [--steps 10].  One JSON line."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402  (weights generator only)
import emrt_b200  # noqa: E402
from emrt_b200 import ops, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--windows", type=int, default=18)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    shapes = synthetic.level_shapes(512)
    Lv = sum(h * w for h, w in shapes)
    B, C = args.windows, 256
    params = O.make_encoder_decoder_params(1234, num_enc=4, num_dec=0)
    enc = emrt_b200.TransformerEncoder(emrt_b200.TransformerEncoderLayer(C, 8, 1024, 0.1, "relu", 3, 6), 4)
    with torch.no_grad():
        for i in range(4):
            sd = enc.layers[i].state_dict()
            for k in sd:
                sd[k].copy_(torch.as_tensor(params[f"encoder.layers.{i}.{k}"]))
    enc = enc.to(dev)
    g = torch.Generator(device="cpu").manual_seed(0)
    srcs = [torch.randn((B, Lv, C), generator=g).mul_(0.5).bfloat16().to(dev) for _ in range(3)]
    pos = torch.randn((1, Lv, C), generator=g).mul_(0.5).bfloat16().to(dev)
    st = torch.tensor(shapes)
    for i in range(args.warmup):
        enc(srcs[i % 3], st, None, pos)
    torch.cuda.synchronize()
    ops.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        enc(srcs[i % 3], st, None, pos)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    rows = B * Lv
    flops = 4 * (2 * rows * 256 * (256 + 432 + 256) + 2 * rows * 2304 * 256 + 2 * 2 * rows * 256 * 1024)
    print(json.dumps({"metric": "TransformerEncoder (4 layers) windows/s", "value": B / (ms * 1e-3), "ms": ms, "windows": B,
                      "gemm_tflops": flops / (ms * 1e-3) / 1e12, "launches": ops.launch_count() // args.steps}))


if __name__ == "__main__":
    main()
