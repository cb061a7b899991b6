"""Batch-consistency / determinism diagnostic of the EncoderDecoder drop-in (GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle.emrt_oracle as O
import emrt_b200
from emrt_b200 import synthetic, ops
dev = torch.device("cuda:0")
st = synthetic.encoder_decoder_state(1234)
m = emrt_b200.EncoderDecoder(hidden_dim=256, dim_feedforward=1024, backbone_num_channels=[512, 1024, 2048], num_feature_levels=3,
                             nhead=8, num_encoder_layers=4, num_decoder_layers=2, num_encoder_points=6, num_decoder_points=6, nclass=6)
with torch.no_grad():
    sd = m.state_dict()
    for k in sd: sd[k].copy_(torch.from_numpy(st[k]))
m = m.to(dev)
rng = np.random.Generator(np.random.PCG64(202))
B, tile = 64, 256
feats = [torch.from_numpy(O.rng_normal(rng, (B, c, tile // s, tile // s), 0.5)).bfloat16().to(dev) for c, s in zip((512, 1024, 2048), (8, 16, 32))]
psp = torch.from_numpy(O.rng_normal(rng, (B, 256, 110), 0.5)).bfloat16().to(dev)
l2 = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()
hs_a, mem_a = m(feats, psp)
hs_b, mem_b = m(feats, psp)
print("same batch twice: mem", l2(mem_a, mem_b), "hs", l2(hs_a, hs_b), "bit-equal", torch.equal(mem_a, mem_b), torch.equal(hs_a, hs_b))
for sub in ([0, 17, 40], [0], list(range(8))):
    hs1, mem1 = m([f[sub] for f in feats], psp[sub])
    print("sub", sub[:4], "mem", l2(mem1, mem_a[sub]), "hs", l2(hs1, hs_a[:, sub]))
# stage by stage on the encoder: layer outputs for the batch vs the sub-batch
sub = [0, 17, 40]
shapes = tuple((tile // s, tile // s) for s in (8, 16, 32))
c = m._constants(shapes, dev, torch.bfloat16)
def enc_in(fs):
    Bn = fs[0].shape[0]
    src = torch.empty((Bn, 1344, 256), dtype=torch.bfloat16, device=dev)
    off = 0
    for l, f in enumerate(fs):
        tok = ops.nchw_to_tokens(f)
        y = ops.linear(tok, c["w"][l], c["b"][l], w_transposed=True)
        ops.groupnorm_tokens_into(y, c["gw"][l], c["gb"][l], src, off, groups=32)
        off += shapes[l][0] * shapes[l][1]
    return src
sa, sb = enc_in(feats), enc_in([f[sub] for f in feats])
print("input_proj+GN", l2(sb, sa[sub]))
xa, xb = sa, sb
ref = emrt_b200.get_reference_points(shapes, device=dev)
for i, layer in enumerate(m.encoder.layers):
    xa = layer(xa, ref, shapes, torch.ones(B, 1344, device=dev), c["pos"])
    xb = layer(xb, ref, shapes, torch.ones(3, 1344, device=dev), c["pos"])
    print("encoder layer", i, l2(xb, xa[sub]))
