#!/usr/bin/env python
"""Encoder-gather time against the spread of the sampling offsets, for the window-staged kernel and for the L1-path kernel
(EMRT_GATHER_NO_WIN=1) on the same inputs: python scripts/gather_sensitivity.py [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import emrt_b200  # noqa: E402
from emrt_b200 import ops, synthetic  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 72
dev = torch.device("cuda:0")
shapes = synthetic.level_shapes(512)
Lv = sum(h * w for h, w in shapes)
g = torch.Generator(device=dev).manual_seed(7)
src = torch.randn((B, Lv, 256), generator=g, device=dev).bfloat16()
pos = torch.randn((1, Lv, 256), generator=g, device=dev).bfloat16()
ref = emrt_b200.get_reference_points(shapes, device=dev)
for sigma in (0.05, 0.1, 0.15, 0.2, 0.3, 0.5):
    m = emrt_b200.MSDeformableAttention(256, 8, 3, 6).to(dev)
    with torch.no_grad():
        for name, arr in synthetic.msda_state(1234, offset_std=sigma).items():
            mod, leaf = name.split(".")
            getattr(getattr(m, mod), leaf).copy_(torch.from_numpy(arr))
    m.requires_grad_(False)
    row = {}
    for tag, env in (("window", None), ("l1", "EMRT_GATHER_NO_WIN")):
        if env:
            os.environ[env] = "1"
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
            if env:
                os.environ.pop(env, None)
        t = [e[0].elapsed_time(e[1]) for (name, dims, e) in ev if name == "msda_gather_fwd"]
        row[tag] = sum(t) / len(t)
    print(f"offset_weight_sigma {sigma:4.2f}: window-staged {row['window']:.3f} ms, L1-path {row['l1']:.3f} ms")
