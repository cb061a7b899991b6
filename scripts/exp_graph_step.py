#!/usr/bin/env python
"""Experiment: the bench step (72 windows, full EncoderDecoder + head tail) replayed from ONE CUDA graph against the same
step launched kernel by kernel — how much of the step is launch gaps.  python scripts/exp_graph_step.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as Bn  # noqa: E402
import emrt_b200  # noqa: E402
from emrt_b200 import ops, _lib as L  # noqa: E402
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


def timeit(fn, iters=20):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


with torch.no_grad():
    eager = timeit(lambda i: hp.step(sets[i % 2], labels[i % 2]))
    graphs = []
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for k in range(2):
            hp.step(sets[k], labels[k])
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=s):
                out = hp.step(sets[k], labels[k])
            graphs.append((gr, out))
        torch.cuda.synchronize()
        replay = timeit(lambda i: graphs[i % 2][0].replay())
print(f"eager {eager:.3f} ms per step, graph replay {replay:.3f} ms per step ({B} windows)")
