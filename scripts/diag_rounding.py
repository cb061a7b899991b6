#!/usr/bin/env python
"""Where do the bf16 kernels differ from the same-rounding-points oracle (oracle.kernel_storage_rounding)?  Stage by stage,
relative L2 against the float64 oracle with the kernels' stores applied (run on the GPU box):
  python scripts/diag_rounding.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402
import emrt_b200  # noqa: E402
from emrt_b200 import ops, _lib as L  # noqa: E402

dev = torch.device("cuda", 0)
l2 = lambda a, b: ((a.detach().double().cpu() - b.double()).norm() / b.double().norm()).item()
r16 = lambda v: torch.as_tensor(v).bfloat16().double()


def load(m, params, prefix=""):
    with torch.no_grad():
        sd = m.state_dict()
        for k in sd:
            sd[k].copy_(torch.as_tensor(params[prefix + k]))
    return m


def both(fn):
    exact = fn()
    with O.kernel_storage_rounding():
        rounded = fn()
    return exact, rounded


STAGES = {}


def rec(name, own):
    """one stage = one kernel (or fused kernel) on the same-rounding oracle's own operands"""
    STAGES[name] = own
    return own


def run(tile=256):
    STAGES.clear()
    shapes = [(tile // 8,) * 2, (tile // 16,) * 2, (tile // 32,) * 2]
    B, C = 2, 256
    _, Lv = O.level_tables(shapes)
    params = O.make_encoder_decoder_params(23, num_enc=2, num_dec=1)
    p64 = {k: (r16(v) if v.ndim >= 2 else torch.as_tensor(v).double()) for k, v in params.items()}
    p64.update({k + "#fp32": torch.as_tensor(v).double() for k, v in params.items() if k.endswith(("sampling_offsets.weight", "attention_weights.weight"))})
    rng = np.random.Generator(np.random.PCG64(24))
    src = torch.from_numpy(O.rng_normal(rng, (B, Lv, C), 0.5)).bfloat16()
    pos = torch.from_numpy(O.rng_normal(rng, (1, Lv, C), 0.5)).bfloat16()
    ref = O.encoder_reference_points(shapes, B).double()
    ones = torch.ones(B, Lv).double()
    pre = "encoder.layers.0."
    sub = O.emrt_oracle._sub

    # 1. MSDA alone, query = src + pos formed outside (no row-bias table), standalone output stored bf16
    q = (src.double() + pos.double()).bfloat16()
    ex, ro = both(lambda: O.msda_forward(sub(p64, pre + "self_attn."), q.double(), ref, src.double(), shapes, ones, dtype=torch.float64))
    m = load(emrt_b200.MSDeformableAttention(C, 8, 3, 6), params, pre + "self_attn.").to(dev).requires_grad_(False)
    ref_d = emrt_b200.get_reference_points(shapes, device=dev)
    with torch.no_grad():
        got = m(q.to(dev), ref_d, src.to(dev), shapes)
    print(f"MSDA alone (window gather): own {l2(got.float(), ro.bfloat16()):.2e}  formats {l2(ro, ex):.2e}  vs exact {l2(got.float(), ex):.2e}")
    # 1b. pieces of it
    with torch.no_grad():
        pk = m.packed_weights()
        v = ops.linear(src.to(dev), pk["wv"], pk["bv"], w_transposed=True)
        with O.kernel_storage_rounding():
            vo, loco, awo = O.msda_intermediates(sub(p64, pre + "self_attn."), q.double(), ref, src.double(), shapes, ones, dtype=torch.float64)
        print(f"  value_proj: own {rec("encoder value_proj", l2(v.float().view(B, Lv, 8, 32), vo)):.2e}")
        off, attn = ops.linear(q.to(dev), pk["wq"], pk["bq"], w_transposed=True, y_dtype=torch.float16, epilogue=L.EPI_MSDA_QPROJ, qproj_group=18)
        norm = torch.tensor([[float(w), float(h)] for h, w in shapes], dtype=torch.float64).reshape(1, 1, 1, 3, 1, 2)
        off_o = (loco - ref.reshape(B, Lv, 1, 3, 1, 2)) * norm
        print(f"  offsets (px): own {rec("query projection: offsets (px)", l2(off.float().view(B, Lv, 8, 3, 6, 2), off_o)):.2e}   softmax weights: own {rec("query projection: softmax weights", l2(attn.float().view(B, Lv, 8, 3, 6), awo)):.2e}")
        # gather on the ORACLE's stored operands: isolates the gather kernel
        for name, mode, vv in (("window", L.LOC_PIXEL_OFFSET | L.VALUE_HEAD_MAJOR | L.QUERY_PIXEL_GRID, vo.permute(0, 2, 1, 3).contiguous()),
                               ("v1 / generic", L.LOC_PIXEL_OFFSET, vo)):
            g = ops.msda_gather_fwd(vv.bfloat16().to(dev), off_o.half().to(dev), awo.half().to(dev), shapes, ref=ref_d, mode=mode)
            with O.kernel_storage_rounding():
                go = O.deformable_attention_core_func(vo, shapes, loco, awo)
            go_exactw = O.deformable_attention_core_func(vo, shapes, loco, awo)
            print(f"  gather ({name}) on the oracle's operands: own {rec("gather (" + name + ")", l2(g.float(), go.bfloat16() if "window" in name else go_exactw.bfloat16())):.2e}   (vs exact-weight gather {l2(g.float(), go_exactw):.2e})")

    # 2. one encoder layer and two
    layer = load(emrt_b200.TransformerEncoderLayer(C, 8, 1024, 0.1, "relu", 3, 6), params, pre).to(dev)
    fn1 = lambda: O.encoder_layer_forward(p64, pre, src.double(), ref, shapes, ones, pos.double().expand(B, -1, -1))
    ex, ro = both(fn1)
    got = layer(src.to(dev), ref_d, shapes, None, pos.to(dev))
    print(f"encoder layer: own {l2(got.float(), ro):.2e}  formats {l2(ro, ex):.2e}  vs exact {l2(got.float(), ex):.2e}")
    # pieces: conv, LN1 output, FFN
    with torch.no_grad():
        pk = layer._packed_weights(torch.bfloat16)
        conv = ops.conv3x3_tokens(src.to(dev), pk["conv_w"], shapes)
        import torch.nn.functional as F
        start, _ = O.level_tables(shapes)
        outs = []
        for l, (h, w) in enumerate(shapes):
            x = src.double()[:, start[l]:start[l] + h * w].permute(0, 2, 1).reshape(B, C, h, w)
            outs.append(F.conv2d(x, p64[f"{pre}conv{l}.0.weight"], None, 1, 1).flatten(2).permute(0, 2, 1))
        print(f"  conv3x3: own {rec("conv3x3", l2(conv.float(), torch.cat(outs, 1).bfloat16())):.2e}")
        x1 = layer.self_attn(src.to(dev), ref_d, src.to(dev), shapes, None, query_pos=pos.to(dev), residual_norm=(src.to(dev), pk["n1w"], pk["n1b"]))
        with O.kernel_storage_rounding():
            s2 = O.msda_forward(sub(p64, pre + "self_attn."), src.double(), ref, src.double(), shapes, ones, dtype=torch.float64, query_pos=pos.double())
            x1o = O.emrt_oracle._store(O.emrt_oracle._ln(src.double() + s2, p64[pre + "norm1.weight"], p64[pre + "norm1.bias"]))
        print(f"  MSDA (+pos row bias) + residual + LN1: own {l2(x1.float(), x1o):.2e}")
        # pieces of that: the row-bias query projection, and the fused output projection + LN on the oracle's gather output
        mm = layer.self_attn
        pkm = mm.packed_weights()
        rowb = mm._query_pos_bias(pos.to(dev), Lv)
        off, attn = ops.linear(src.to(dev), pkm["wq"], None, w_transposed=True, y_dtype=torch.float16, epilogue=L.EPI_MSDA_QPROJ,
                               qproj_group=18, row_bias=rowb, row_bias_period=Lv)
        with O.kernel_storage_rounding():
            vo2, loco2, awo2 = O.msda_intermediates(sub(p64, pre + "self_attn."), src.double(), ref, src.double(), shapes, ones,
                                                    dtype=torch.float64, query_pos=pos.double())
            go2 = O.emrt_oracle._store(O.deformable_attention_core_func(vo2, shapes, loco2, awo2))
        off_o2 = (loco2 - ref.reshape(B, Lv, 1, 3, 1, 2)) * norm
        print(f"    row-bias offsets (px): own {rec("row-bias query projection: offsets (px)", l2(off.float().view(B, Lv, 8, 3, 6, 2), off_o2)):.2e} max abs {(off.float().view(B, Lv, 8, 3, 6, 2).cpu().double() - off_o2).abs().max().item():.2e}"
              f"   softmax weights: own {rec("row-bias query projection: softmax weights", l2(attn.float().view(B, Lv, 8, 3, 6), awo2)):.2e}")
        y = ops.linear(go2.bfloat16().to(dev), pkm["wo"], pkm["bo"], w_transposed=True, epilogue=L.EPI_RESIDUAL_LN, residual=src.to(dev),
                       ln_gamma=pk["n1w"], ln_beta=pk["n1b"])
        s2o = go2 @ p64[pre + "self_attn.output_proj.weight"] + p64[pre + "self_attn.output_proj.bias"]
        yo = O.emrt_oracle._ln(src.double() + s2o, p64[pre + "norm1.weight"], p64[pre + "norm1.bias"]).bfloat16()
        print(f"    fused output_proj + residual + LN1 on the oracle's gather output: own {rec("output_proj + residual + LN1 (fused)", l2(y.float(), yo)):.2e}")
        y2 = ops.linear(go2.bfloat16().to(dev), pkm["wo"], pkm["bo"], w_transposed=True)
        print(f"    plain output_proj: own {rec("output_proj", l2(y2.float(), s2o.bfloat16())):.2e}")
        h = ops.linear(x1o.bfloat16().to(dev), pk["w1"], pk["b1"], w_transposed=True, epilogue=L.EPI_RELU)
        ho = F.relu(x1o @ p64[pre + "linear1.weight"] + p64[pre + "linear1.bias"]).bfloat16()
        print(f"  FFN1 on the oracle's LN1 output: own {rec("FFN linear1 + ReLU", l2(h.float(), ho)):.2e}")
        f = ops.linear(ho.to(dev), pk["w2"], pk["b2"], w_transposed=True)
        fo = (ho.double() @ p64[pre + "linear2.weight"] + p64[pre + "linear2.bias"]).bfloat16()
        print(f"  FFN2: own {rec("FFN linear2", l2(f.float(), fo)):.2e}")
        gn = ops.groupnorm_stats(torch.cat(outs, 1).bfloat16().to(dev), shapes, groups=32)
        y = ops.residual_layernorm_gn(fo.to(dev), x1o.bfloat16().to(dev), pk["n2w"], pk["n2b"], torch.cat(outs, 1).bfloat16().to(dev), src.to(dev),
                                      gn, pk["gn_w"], pk["gn_b"], shapes, groups=32)
        br = []
        for l, (h_, w_) in enumerate(shapes):
            x = src.double()[:, start[l]:start[l] + h_ * w_].permute(0, 2, 1).reshape(B, C, h_, w_)
            yy = F.group_norm(outs[l].bfloat16().double().permute(0, 2, 1).reshape(B, C, h_, w_), 32, p64[f"{pre}conv{l}.1.weight"], p64[f"{pre}conv{l}.1.bias"], 1e-5)
            br.append((F.gelu(yy) + x).flatten(2).permute(0, 2, 1))
        yo = (O.emrt_oracle._ln(x1o + fo.double(), p64[pre + "norm2.weight"], p64[pre + "norm2.bias"]) + torch.cat(br, 1)).bfloat16()
        yf = ops.linear(ho.to(dev), pk["w2"], pk["b2"], w_transposed=True, epilogue=L.EPI_RESIDUAL_LN, residual=x1o.bfloat16().to(dev),
                        ln_gamma=pk["n2w"], ln_beta=pk["n2b"],
                        gn_branch=dict(conv=torch.cat(outs, 1).bfloat16().to(dev), skip=src.to(dev), stats=gn, gamma=pk["gn_w"],
                                       beta=pk["gn_b"], shapes=shapes))
        yfo = (O.emrt_oracle._ln(x1o + ho.double() @ p64[pre + "linear2.weight"] + p64[pre + "linear2.bias"], p64[pre + "norm2.weight"],
                                 p64[pre + "norm2.bias"]) + torch.cat(br, 1)).bfloat16()
        print(f"  linear2 + residual + LN2 + GroupNorm + GELU + skip (fused): own {rec('FFN linear2 + residual + LN2 + conv branch (fused)', l2(yf.float(), yfo)):.2e}")
        print(f"  LN2 + GroupNorm + GELU + skip: own {rec("LN2 + GroupNorm + GELU + skip", l2(y.float(), yo)):.2e}")
        # the whole FFN in one kernel (emrt_ffn_fused_fwd): hidden chunk rounded to bf16 on chip, x + linear2 parked as bf16
        yff = ops.ffn_fused(x1o.bfloat16().to(dev), pk["w1"], pk["b1"], pk["w2"], pk["b2"], pk["n2w"], pk["n2b"],
                            gn_branch=dict(conv=torch.cat(outs, 1).bfloat16().to(dev), skip=src.to(dev), stats=gn, gamma=pk["gn_w"],
                                           beta=pk["gn_b"], shapes=shapes))
        pre_ln = (x1o + ho.double() @ p64[pre + "linear2.weight"] + p64[pre + "linear2.bias"]).bfloat16().double()
        yffo = (O.emrt_oracle._ln(pre_ln, p64[pre + "norm2.weight"], p64[pre + "norm2.bias"]) + torch.cat(br, 1)).bfloat16()
        print(f"  whole FFN + LN2 + conv branch in one kernel: own {rec('FFN (linear1 + ReLU + linear2) + residual + LN2 + conv branch (one kernel)', l2(yff.float(), yffo)):.2e}")

    # 3. decoder layer, stage by stage on the oracle's operands
    print("decoder layer")
    E = O.emrt_oracle
    Nq = 110
    dpre = "decoder.layers.0."
    tgt = torch.from_numpy(O.rng_normal(rng, (B, Nq, C), 0.5)).bfloat16()
    qpos = torch.from_numpy(O.rng_normal(rng, (1, Nq, C), 1.0)).bfloat16()
    mem = torch.from_numpy(O.rng_normal(rng, (B, Lv, C), 1.0)).bfloat16()
    refd = torch.from_numpy(np.repeat(rng.uniform(0.1, 0.9, size=(1, Nq, 1, 2)).astype(np.float32), 3, axis=2))
    dl = load(emrt_b200.TransformerDecoderLayer(C, 8, 1024, 0.1, "relu", 3, 6), params, dpre).to(dev)
    with torch.no_grad():
        got = dl(tgt.to(dev), refd.to(dev), mem.to(dev), shapes, None, qpos.to(dev))
        with O.kernel_storage_rounding():
            want = E.decoder_layer_forward(p64, dpre, tgt.double(), refd.double().expand(B, -1, -1, -1), mem.double(), shapes, ones,
                                           qpos.double().expand(B, -1, -1))
        print(f"  whole layer: own {l2(got.float(), want):.2e}")
        pk = dl._packed_weights(torch.bfloat16)
        x2 = ops.cyclic_rows_cached(qpos.to(dev))
        qk = ops.linear(tgt.to(dev), pk["w_qk"], pk["b_qk"], w_transposed=True, x2=x2, x2_period=Nq)
        v = ops.linear(tgt.to(dev), pk["w_v"], pk["b_v"], w_transposed=True)
        Wi, bi = p64[dpre + "self_attn.in_proj_weight"], p64[dpre + "self_attn.in_proj_bias"]
        qin = tgt.double() + qpos.double()
        qko = (qin @ Wi[:, :2 * C] + bi[:2 * C]).bfloat16()
        vo = (tgt.double() @ Wi[:, 2 * C:] + bi[2 * C:]).bfloat16()
        print(f"  qk projection (pos folded as k-blocks): own {rec("decoder q/k projection (pos folded)", l2(qk.float(), qko)):.2e}   v: own {rec("decoder v projection", l2(v.float(), vo)):.2e}")
        att = ops.mha_small(qko.to(dev)[..., :C], qko.to(dev)[..., C:], vo.to(dev), 8, 32 ** -0.5)
        hd = lambda t: t.double().reshape(B, Nq, 8, 32).permute(0, 2, 1, 3)
        w = torch.softmax(hd(qko[..., :C]) @ hd(qko[..., C:]).transpose(-1, -2) * 32 ** -0.5, -1)
        atto = (w @ hd(vo)).permute(0, 2, 1, 3).reshape(B, Nq, C).bfloat16()
        print(f"  mha_small on the oracle's q/k/v: own {rec("decoder self-attention core", l2(att.float(), atto)):.2e}   (max softmax weight {w.max().item():.3f})")
        t1 = ops.linear(atto.to(dev), pk["w_o"], pk["b_o"], w_transposed=True, epilogue=L.EPI_RESIDUAL_LN, residual=tgt.to(dev),
                        ln_gamma=pk["n1w"], ln_beta=pk["n1b"])
        t1o = E._ln(tgt.double() + atto.double() @ p64[dpre + "self_attn.out_proj.weight"] + p64[dpre + "self_attn.out_proj.bias"],
                    p64[dpre + "norm1.weight"], p64[dpre + "norm1.bias"]).bfloat16()
        print(f"  out_proj + residual + LN1: own {rec("decoder out_proj + residual + LN1 (fused)", l2(t1.float(), t1o)):.2e}")
        t2 = dl.cross_attn(t1o.to(dev), refd.to(dev), mem.to(dev), shapes, None, query_pos=qpos.to(dev), residual_norm=(t1o.to(dev), pk["n2w"], pk["n2b"]))
        with O.kernel_storage_rounding():
            s2 = O.msda_forward(sub(p64, dpre + "cross_attn."), t1o.double(), refd.double().expand(B, -1, -1, -1), mem.double(), shapes, ones,
                                dtype=torch.float64, query_pos=qpos.double())
            t2o = E._store(E._ln(t1o.double() + s2, p64[dpre + "norm2.weight"], p64[dpre + "norm2.bias"]))
        print(f"  cross attention + residual + LN2: own {rec("decoder cross attention (4 kernels) + residual + LN2", l2(t2.float(), t2o)):.2e}")
        hh = ops.linear(t2o.bfloat16().to(dev), pk["w1"], pk["b1"], w_transposed=True, epilogue=L.EPI_RELU)
        hho = torch.relu(t2o @ p64[dpre + "linear1.weight"] + p64[dpre + "linear1.bias"]).bfloat16()
        t3 = ops.linear(hho.to(dev), pk["w2"], pk["b2"], w_transposed=True, epilogue=L.EPI_RESIDUAL_LN, residual=t2o.bfloat16().to(dev),
                        ln_gamma=pk["n3w"], ln_beta=pk["n3b"])
        t3o = E._ln(t2o + hho.double() @ p64[dpre + "linear2.weight"] + p64[dpre + "linear2.bias"], p64[dpre + "norm3.weight"], p64[dpre + "norm3.bias"]).bfloat16()
        print(f"  FFN1: own {rec("decoder FFN linear1 + ReLU", l2(hh.float(), hho)):.2e}   FFN2 + residual + LN3: own {rec("decoder FFN linear2 + residual + LN3 (fused)", l2(t3.float(), t3o)):.2e}")
    return dict(STAGES)


if __name__ == "__main__":
    out = run(int(sys.argv[1]) if len(sys.argv) > 1 else 256)
    print("\nstages (kernels vs same-rounding oracle, each on the oracle's operands):")
    for k, v in out.items():
        print(f"  {k:58s} {v:.2e}")
