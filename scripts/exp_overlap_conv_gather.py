#!/usr/bin/env python
"""Experiment: the encoder layer's 3x3 conv (tensor-pipe bound, power-limited when it owns all 148 SMs) on K SMs of a side
stream WHILE the sampling gather (shared-memory-pipe bound) runs on the rest, against the two back to back.
python scripts/exp_overlap_conv_gather.py [K ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import emrt_b200  # noqa: E402
from emrt_b200 import ops, synthetic, _lib as L  # noqa: E402

B = 72
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
conv_w = ops.pack_conv3x3_weights([torch.randn((256, 256, 3, 3), generator=g, device=dev) * 0.02 for _ in shapes], torch.bfloat16)

# capture the gather's arguments from the launch-by-launch forward
captured = {}
real_gather = ops.msda_gather_fwd


def spy(*a, **k):
    captured["a"], captured["k"] = a, k
    return real_gather(*a, **k)


m.gemm_impl = L.IMPL_TCGEN05
ops.msda_gather_fwd = spy
with torch.no_grad():
    m(src, ref, src, shapes, query_pos=pos)
ops.msda_gather_fwd = real_gather
torch.cuda.synchronize()
ga, gk = captured["a"], captured["k"]
gather = lambda: real_gather(*ga, **gk)
conv = lambda k=0: ops.conv3x3_tokens_stats(src, conv_w, shapes, max_ctas=k)


def timed(fn, iters=10):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


side = torch.cuda.Stream(priority=-1)
main = torch.cuda.current_stream()


def both(k):
    def fn():
        fork = torch.cuda.Event(); fork.record(main)
        with torch.cuda.stream(side):
            side.wait_event(fork)
            conv(k)
            done = torch.cuda.Event(); done.record(side)
        gather()
        main.wait_event(done)
    return fn


with torch.no_grad():
    tg, tc = timed(gather), timed(conv)
    print(f"gather alone {tg:.0f} us, conv alone (148 SMs) {tc:.0f} us, back to back {timed(lambda: (gather(), conv())):.0f} us")
    for k in [int(x) for x in sys.argv[1:]] or [24, 32, 40, 48, 64]:
        print(f"conv on {k} SMs of a side stream || gather: {timed(both(k)):.0f} us")
