#!/usr/bin/env python
"""Experiment: FFN1 -> FFN2 (+ norm2 + conv branch) of one encoder layer at the bench geometry, run chunk by chunk over the
batch so that a chunk's hidden tensor [rows, 1024] can stay in the 126 MB L2 between the two GEMMs (one h buffer, reused by
every chunk).  Replayed from a CUDA graph, so host launch overhead is not in the numbers.
python scripts/exp_ffn_chunked.py [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import emrt_b200  # noqa: E402
from emrt_b200 import ops, synthetic, _lib as L  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 72
dev = torch.device("cuda", 0)
shapes = synthetic.level_shapes(512)
Lv = sum(h * w for h, w in shapes)
g = torch.Generator(device=dev).manual_seed(7)
src = torch.randn((B, Lv, 256), generator=g, device=dev).bfloat16()
x = torch.randn((B, Lv, 256), generator=g, device=dev).bfloat16()
conv = torch.randn((B, Lv, 256), generator=g, device=dev).bfloat16()
layer = emrt_b200.TransformerEncoderLayer(256, 8, 1024, 0.0, "relu", 3, 6).to(dev).requires_grad_(False)
with torch.no_grad():
    st = synthetic.encoder_layer_state(1234)
    sd = layer.state_dict()
    for k in sd:
        sd[k].copy_(torch.from_numpy(st[k]))
    lp = layer._packed_weights(torch.bfloat16)
    stats = ops.groupnorm_stats(conv, shapes, groups=32)
    out = torch.empty_like(x)

    def run(chunk):
        h = torch.empty((chunk, Lv, 1024), dtype=torch.bfloat16, device=dev)
        for b0 in range(0, B, chunk):
            b1 = min(B, b0 + chunk)
            hh = h[: b1 - b0]
            ops.linear(x[b0:b1], lp["w1"], lp["b1"], w_transposed=True, epilogue=L.EPI_RELU, out=hh)
            ops.linear(hh, lp["w2"], lp["b2"], w_transposed=True, epilogue=L.EPI_RESIDUAL_LN, residual=x[b0:b1], ln_gamma=lp["n2w"],
                       ln_beta=lp["n2b"], out=out[b0:b1],
                       gn_branch=dict(conv=conv[b0:b1], skip=src[b0:b1], stats=stats[b0 * 3 * 32 * 2:], gamma=lp["gn_w"], beta=lp["gn_b"], shapes=shapes))
        return h

    for chunk in (72, 36, 18, 12, 9, 6, 4, 3, 2):
        if chunk > B:
            continue
        keep = run(chunk)
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=s):
                keep = run(chunk)
            gr.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(10):
                gr.replay()
            e1.record(s)
            torch.cuda.synchronize()
        print(f"chunk {chunk:3d} windows  h buffer {chunk * Lv * 2048 / 1e6:7.1f} MB  FFN1 + FFN2/LN/branch: {e0.elapsed_time(e1) / 10 * 1e3:8.1f} us", flush=True)
