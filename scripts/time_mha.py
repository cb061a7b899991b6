#!/usr/bin/env python
"""Times the decoder self-attention core (emrt_mha_small) at the bench geometry: 72 windows x 110 queries x 8 heads."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from emrt_b200 import ops
dev = torch.device("cuda", 0)
B, Lq, C = 72, 110, 256
g = torch.Generator(device=dev).manual_seed(1)
qk = torch.randn((B, Lq, 2 * C), generator=g, device=dev).bfloat16()
v = torch.randn((B, Lq, C), generator=g, device=dev).bfloat16()
fn = lambda: ops.mha_small(qk[..., :C], qk[..., C:], v, 8, 32 ** -0.5)
fn(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): fn()
e1.record(); torch.cuda.synchronize()
print("mha_small avg us", e0.elapsed_time(e1) / 20 * 1e3)
