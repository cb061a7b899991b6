#!/usr/bin/env python
"""Two encoder MSDeformableAttention calls at the bench geometry (B windows of 512x512, bf16) — the target of the ncu
captures: bench.py runs it under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` to measure `roofline.traffic`
in the run itself, scripts/gpu_check.sh under `ncu --set full` for the committed capture."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import emrt_b200  # noqa: E402
from emrt_b200 import synthetic  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 72
dev = torch.device("cuda", 0)
shapes = synthetic.level_shapes(512)
Lv = sum(h * w for h, w in shapes)
g = torch.Generator(device=dev).manual_seed(7)
src = torch.randn((B, Lv, 256), generator=g, device=dev).bfloat16()
pos = torch.randn((1, Lv, 256), generator=g, device=dev).bfloat16()
ref = emrt_b200.get_reference_points(shapes, device=dev)
m = emrt_b200.MSDeformableAttention(256, 8, 3, 6).to(dev).requires_grad_(False)
with torch.no_grad():
    for name, arr in synthetic.msda_state(1234).items():
        mod, leaf = name.split(".")
        getattr(getattr(m, mod), leaf).copy_(torch.from_numpy(arr))
    for _ in range(2):
        m(src, ref, src, shapes, query_pos=pos)
torch.cuda.synchronize()
