#!/usr/bin/env python
"""One kernel of the bench step at the bench geometry (B windows of 512x512, bf16), launched three times — the target of
single-kernel `ncu --set full` captures (scripts/gpu_check.sh):  python scripts/op_once.py qproj|ffn2|ffn1|value [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import emrt_b200  # noqa: E402
from emrt_b200 import ops, synthetic, _lib as L  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "qproj"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 72
dev = torch.device("cuda", 0)
shapes = synthetic.level_shapes(512)
Lv = sum(h * w for h, w in shapes)
g = torch.Generator(device=dev).manual_seed(7)
src = torch.randn((B, Lv, 256), generator=g, device=dev).bfloat16()
pos = torch.randn((1, Lv, 256), generator=g, device=dev).bfloat16()
layer = emrt_b200.TransformerEncoderLayer(256, 8, 1024, 0.0, "relu", 3, 6).to(dev).requires_grad_(False)
with torch.no_grad():
    st = synthetic.encoder_layer_state(1234)
    sd = layer.state_dict()
    for k in sd:
        sd[k].copy_(torch.from_numpy(st[k]))
    m = layer.self_attn
    pk, lp = m.packed_weights(), layer._packed_weights(torch.bfloat16)
    if what == "qproj":
        rowb = m._query_pos_bias(pos, Lv)
        fn = lambda: ops.linear(src, pk["wq"], None, w_transposed=True, y_dtype=torch.float16, epilogue=L.EPI_MSDA_QPROJ,
                                qproj_group=18, row_bias=rowb, row_bias_period=Lv)
    elif what == "value":
        fn = lambda: ops.linear(src, pk["wv"], pk["bv"], w_transposed=True, epilogue=L.EPI_HEAD_MAJOR, hm_rows=Lv, hm_D=32)
    elif what == "ffn1":
        fn = lambda: ops.linear(src, lp["w1"], lp["b1"], w_transposed=True, epilogue=L.EPI_RELU)
    elif what == "ffn2":
        h = torch.randn((B, Lv, 1024), generator=g, device=dev).bfloat16()
        conv = torch.randn((B, Lv, 256), generator=g, device=dev).bfloat16()
        stats = ops.groupnorm_stats(conv, shapes, groups=32)
        fn = lambda: ops.linear(h, lp["w2"], lp["b2"], w_transposed=True, epilogue=L.EPI_RESIDUAL_LN, residual=src, ln_gamma=lp["n2w"],
                                ln_beta=lp["n2b"], gn_branch=dict(conv=conv, skip=src, stats=stats, gamma=lp["gn_w"], beta=lp["gn_b"], shapes=shapes))
    elif what == "conv":
        fn = lambda: ops.conv3x3_tokens(src, lp["conv_w"], shapes)
    elif what == "convs":
        fn = lambda: ops.conv3x3_tokens_stats(src, lp["conv_w"], shapes, max_ctas=int(os.environ.get("EMRT_CONV_MAX_CTAS", "0")))
    elif what == "ffn":
        conv = torch.randn((B, Lv, 256), generator=g, device=dev).bfloat16()
        stats = ops.groupnorm_stats(conv, shapes, groups=32)
        x_in = torch.randn((B, Lv, 256), generator=g, device=dev).bfloat16()
        out = torch.empty_like(src)
        fn = lambda: ops.ffn_fused(x_in, lp["w1"], lp["b1"], lp["w2"], lp["b2"], lp["n2w"], lp["n2b"], out=out,
                                   gn_branch=dict(conv=conv, skip=src, stats=stats, gamma=lp["gn_w"], beta=lp["gn_b"], shapes=shapes))
    else:
        raise SystemExit("unknown op " + what)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = int(os.environ.get("EMRT_OP_ITERS", "3"))
    fn()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(what, "avg ms", e0.elapsed_time(e1) / iters)
