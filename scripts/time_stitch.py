#!/usr/bin/env python
"""Times emrt_stitch_argmax_fused at the bench geometry (8 images of 1024 x 1024, nine 512 x 512 windows each at stride 384, 7 classes, bf16
half-resolution logits): python scripts/time_stitch.py [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import emrt_b200  # noqa: E402
from emrt_b200 import ops  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda:0")
n_img, H, W, hc, wc, nc = 8, 1024, 1024, 512, 512, 7
plan, _, _ = emrt_b200.plan_windows([(H, W)] * n_img, (wc, hc), (384, 384))
g = torch.Generator(device="cpu").manual_seed(5)
half = [torch.randn((len(plan), nc, hc // 2, wc // 2), generator=g).to(dev).bfloat16() for _ in range(3)]   # 3 x 66 MB > L2
t = lambda k: torch.tensor([p[k] for p in plan], dtype=torch.int32, device=dev)
wi, wy, wx = t(0), t(1), t(2)
for s in range(3):
    lab = ops.stitch_argmax_fused(half[s], wi, wy, wx, n_img, H, W, label_dtype=torch.uint8)[0]
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(iters):
    lab = ops.stitch_argmax_fused(half[i % 3], wi, wy, wx, n_img, H, W, label_dtype=torch.uint8)[0]
e1.record()
torch.cuda.synchronize()
print(f"stitch {len(plan)} windows: {e0.elapsed_time(e1) / iters * 1e3:.1f} us per call, label checksum {int(lab.long().sum())}")
