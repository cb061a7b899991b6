"""Stage-by-stage bf16 error of the decoder layer against the float64 oracle (diagnostic, GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F
import oracle.emrt_oracle as O
import emrt_b200
from emrt_b200 import ops, _lib as L

dev = torch.device("cuda:0")
l2 = lambda got, want: ((got.double().cpu() - want).norm() / want.norm()).item()
mx = lambda got, want: ((got.double().cpu() - want).abs().max() / want.abs().max()).item()


def load(module, params, prefix=""):
    with torch.no_grad():
        sd = module.state_dict()
        for k in sd:
            sd[k].copy_(torch.as_tensor(params[prefix + k]))
    return module


params = O.make_encoder_decoder_params(43, num_enc=4, num_dec=2)
rng = np.random.Generator(np.random.PCG64(44))
chans = (512, 1024, 2048)
feats = [torch.from_numpy(O.rng_normal(rng, (2, c, 256 // s, 256 // s), 0.5)).bfloat16() for c, s in zip(chans, (8, 16, 32))]
psp = torch.from_numpy(O.rng_normal(rng, (2, 256, 110), 0.5)).bfloat16()
r16 = lambda v: torch.as_tensor(v).bfloat16().double()
keep = lambda k: k.endswith("embed.weight") or k == "reference_points.weight"
p64 = {k: (r16(v) if v.ndim >= 2 and not keep(k) else torch.as_tensor(v).double()) for k, v in params.items()}
whs, wmem, shapes = O.encoder_decoder_forward(p64, [f.double() for f in feats], psp.double(), num_enc=4, num_dec=2)
m = load(emrt_b200.EncoderDecoder(110, "sine", False, chans, 3, 6, 6, 6, 256, 8, 4, 2, 1024), params).to(dev)
hs, mem = m([f.to(dev) for f in feats], psp.to(dev))
print("memory l2", l2(mem.float(), wmem), "hs l2", l2(hs.float(), whs))

# ---- decoder layer 0, stage by stage, fed with the ORACLE's memory (rounded to bf16) -------------------------------
bs = 2
pre = "decoder.layers.0."
p = p64
qe = p["query_pos_embed.weight"][None].expand(bs, -1, -1)
rp = torch.sigmoid(qe @ p["reference_points.weight"] + p["reference_points.bias"])[:, :, None, :].expand(-1, -1, 3, -1)
tgt = psp.double().permute(0, 2, 1)
mask = torch.ones(bs, wmem.shape[1]).double()
# oracle stages
q = tgt + qe
o_att = O.multi_head_attention(p, pre + "self_attn.", q, q, tgt)
o_t1 = O._ln(tgt + o_att, p[pre + "norm1.weight"], p[pre + "norm1.bias"])
sub = O._sub(p, pre + "cross_attn.")
inter = O.msda_intermediates(sub, o_t1 + qe, rp, wmem, shapes, mask, dtype=torch.float64)
o_ca = O.msda_forward(sub, o_t1 + qe, rp, wmem, shapes, mask, dtype=torch.float64)
o_t2 = O._ln(o_t1 + o_ca, p[pre + "norm2.weight"], p[pre + "norm2.bias"])
o_ffn = F.relu(o_t2 @ p[pre + "linear1.weight"] + p[pre + "linear1.bias"]) @ p[pre + "linear2.weight"] + p[pre + "linear2.bias"]
o_t3 = O._ln(o_t2 + o_ffn, p[pre + "norm3.weight"], p[pre + "norm3.bias"])

layer = m.decoder.layers[0]
c = m._constants(tuple(shapes), dev, torch.bfloat16)
pk = layer._packed_weights(torch.bfloat16)
C_, M = 256, 8
lin = lambda x, w, b, **kw: ops.linear(x, pk[w], pk[b], w_transposed=True, impl=L.IMPL_AUTO, **kw)
pos = c["qpos"]
t = ops.nchw_to_tokens(psp.to(dev))
print("tgt tokens", l2(t.float(), tgt))
qk = lin(ops.add_bcast(t, pos), "w_qk", "b_qk")
v = lin(t, "w_v", "b_v")
att = ops.mha_small(qk[..., :C_], qk[..., C_:], v, M, float(C_ // M) ** -0.5)
t2 = lin(att, "w_o", "b_o")
print("self-attn out l2 %.4g max %.4g" % (l2(t2.float(), o_att), mx(t2.float(), o_att)))
t1 = ops.residual_layernorm(t2, t, pk["n1w"], pk["n1b"])
print("after norm1   l2 %.4g max %.4g" % (l2(t1.float(), o_t1), mx(t1.float(), o_t1)))
for name, memory in (("oracle memory", wmem.bfloat16().to(dev)), ("our memory", mem)):
    ca = layer.cross_attn(ops.add_bcast(t1, pos), c["ref_dec"], memory, tuple(shapes), torch.ones(bs, wmem.shape[1], device=dev))
    print(f"[{name}] cross-attn out l2 %.4g max %.4g" % (l2(ca.float(), o_ca), mx(ca.float(), o_ca)))
    # same with the oracle's norm1 output as the query (isolates the MSDA module)
    ca2 = layer.cross_attn(ops.add_bcast(o_t1.bfloat16().to(dev), pos), c["ref_dec"], memory, tuple(shapes),
                           torch.ones(bs, wmem.shape[1], device=dev))
    print(f"[{name}] cross-attn out (oracle query) l2 %.4g" % l2(ca2.float(), o_ca))
    tt2 = ops.residual_layernorm(ca, t1, pk["n2w"], pk["n2b"])
    print(f"[{name}] after norm2 l2 %.4g" % l2(tt2.float(), o_t2))
    h = lin(tt2, "w1", "b1", epilogue=L.EPI_RELU)
    f = lin(h, "w2", "b2")
    print(f"[{name}] ffn out l2 %.4g" % l2(f.float(), o_ffn))
    t3 = ops.residual_layernorm(f, tt2, pk["n3w"], pk["n3b"])
    print(f"[{name}] after norm3 l2 %.4g" % l2(t3.float(), o_t3))
# sizes of things
print("magnitudes: self-attn out rms %.3g, tgt rms %.3g, cross-attn rms %.3g, ffn rms %.3g" %
      (o_att.pow(2).mean().sqrt(), tgt.pow(2).mean().sqrt(), o_ca.pow(2).mean().sqrt(), o_ffn.pow(2).mean().sqrt()))
loc = inter["sampling_locations"] if isinstance(inter, dict) and "sampling_locations" in inter else None
if loc is not None:
    print("loc range", float(loc.min()), float(loc.max()))

# ---- inside the cross-attention module: stage by stage ----------------------------------------------------------------
print("---- MSDA internals (oracle query, oracle memory) ----")
ms = layer.cross_attn
pkm = ms.packed_weights()
query64 = o_t1 + qe
qd = query64.bfloat16().to(dev)
memd = wmem.bfloat16().to(dev)
ov, oloc, oaw = O.msda_intermediates(sub, query64, rp, wmem, shapes, mask, dtype=torch.float64)
Mh, P, D = 8, 6, 32
v_hm = ops.linear(memd, pkm["wv"], pkm["bv"], w_transposed=True, epilogue=L.EPI_ROW_MASK | L.EPI_HEAD_MAJOR,
                  row_scale=torch.ones(bs * wmem.shape[1], device=dev), hm_rows=wmem.shape[1], hm_D=D)
v_pm = v_hm.view(bs, Mh, -1, D).permute(0, 2, 1, 3)
print("value_proj l2 %.4g" % l2(v_pm.float(), ov))
off_px, attn = ops.linear(qd, pkm["wq"], pkm["bq"], w_transposed=True, y_dtype=torch.float16, epilogue=L.EPI_MSDA_QPROJ,
                          qproj_group=18)
off_px = off_px.view(bs, 110, Mh, 3, P, 2)
attn = attn.view(bs, 110, Mh, 3, P)
norm = torch.tensor([[float(w), float(h)] for h, w in shapes]).double().reshape(1, 1, 1, 3, 1, 2)
refb = rp.reshape(bs, 110, 1, 3, 1, 2)
want_off = (oloc - refb) * norm
print("offsets(px) l2 %.4g  max abs err %.4g" % (l2(off_px.float(), want_off), (off_px.double().cpu() - want_off).abs().max()))
print("attn l2 %.4g  max abs err %.4g" % (l2(attn.float(), oaw), (attn.double().cpu() - oaw).abs().max()))
# gather kernel alone: our kernel on OUR v / offsets / attn against the oracle gather evaluated on the same numbers
loc_ours = refb + off_px.double().cpu() / norm
g_want = O.deformable_attention_core_func(v_pm.double().cpu(), shapes, loc_ours, attn.double().cpu())
g_got = ops.msda_gather_fwd(v_hm.view(bs, Mh, -1, D), off_px, attn, tuple(shapes), ref=c["ref_dec"],
                            mode=L.LOC_PIXEL_OFFSET | L.VALUE_HEAD_MAJOR)
print("gather kernel alone l2 %.4g" % l2(g_got.float(), g_want))
g_exact = O.deformable_attention_core_func(ov, shapes, oloc, oaw)
print("gather out vs exact l2 %.4g" % l2(g_got.float(), g_exact))
rp_ours = c["ref_dec"].double().cpu()
print("ref points max abs err %.4g" % (rp_ours - rp[:1]).abs().max())
