#!/usr/bin/env python
"""How much of the bench step is idle time between kernels: CUPTI kernel start / end timestamps (torch.profiler) of a few
eager steps -> per step: sum of kernel durations, sum of the gaps between consecutive kernels, the largest gaps and what
follows them.  python scripts/exp_step_gaps.py"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as Bn  # noqa: E402
from emrt_b200.hotpath import HotPath  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
n_img = 8
B = n_img * Bn.WINDOWS_PER_IMAGE
hp = HotPath(dev, Bn.TILE, Bn.NC, mode="full")
plan, H, W = Bn.window_tables(n_img)
ti = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)
win = dict(win_img=ti([p[0] for p in plan]), win_y0=ti([p[1] for p in plan]), win_x0=ti([p[2] for p in plan]), n_img=n_img, H=H, W=W)
g = torch.Generator(device=dev).manual_seed(3)
sets = []
for s in range(2):
    feats = [(torch.randn((B, c, Bn.TILE // st, Bn.TILE // st), generator=g, device=dev) * 0.5).bfloat16() for c, st in zip(Bn.FEAT_CH, (8, 16, 32))]
    psp = (torch.randn((B, hp.C, hp.num_queries), generator=g, device=dev) * 0.5).bfloat16()
    hl = torch.randn((B, Bn.NC, Bn.TILE // 2, Bn.TILE // 2), generator=g, device=dev).bfloat16()
    sets.append(dict(feats=feats, psp=psp, half_logits=hl, **win))
labels = [torch.empty((n_img, 1, H, W), dtype=torch.uint8, device=dev) for _ in range(2)]
STEPS = 4
with torch.no_grad():
    for i in range(3):
        hp.step(sets[i % 2], labels[i % 2])
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(STEPS):
            hp.step(sets[i % 2], labels[i % 2])
        torch.cuda.synchronize()
ev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start),
            key=lambda e: e.time_range.start)
busy = sum(e.time_range.end - e.time_range.start for e in ev)
span = ev[-1].time_range.end - ev[0].time_range.start
gaps = [(ev[i + 1].time_range.start - ev[i].time_range.end, ev[i].name[:50], ev[i + 1].name[:50]) for i in range(len(ev) - 1)]
gap_sum = sum(max(g_[0], 0) for g_ in gaps)
print(f"{len(ev)} kernels in {STEPS} steps: span {span / STEPS / 1e3:.3f} ms per step, kernels {busy / STEPS / 1e3:.3f} ms, "
      f"gaps {gap_sum / STEPS / 1e3:.3f} ms ({100 * gap_sum / span:.1f} %), mean gap {gap_sum / len(gaps):.2f} us")
for gp, a, b in sorted(gaps, reverse=True)[:12]:
    print(f"  {gp:7.1f} us between {a} -> {b}")
