import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests/golden")
import numpy as np, torch
import oracle as O
import emrt_b200
import make_reference_vectors as G
dev = torch.device("cuda", 0)
g = np.load("/root/repo/tests/golden/ref_encdec_masked.npz")
c = G.encdec_inputs(64, 2, 70, 2, 1)
m = emrt_b200.EncoderDecoder(hidden_dim=256, dim_feedforward=1024, backbone_num_channels=[512, 1024, 2048], num_feature_levels=3, nhead=8,
                             num_encoder_layers=2, num_decoder_layers=1, num_encoder_points=6, num_decoder_points=6, nclass=6)
with torch.no_grad():
    sd = m.state_dict()
    for k in sd: sd[k].copy_(torch.as_tensor(c["params"][k]))
m = m.to(dev)
d = lambda a: torch.from_numpy(a).to(dev)
hs, mem = m([d(f) for f in c["feats"]], d(c["psp"]), d(g["src_mask"]))
err = (mem.cpu().numpy() - g["memory"])
scale = np.abs(g["memory"]).max()
print("max err per image:", np.abs(err).max(axis=(1, 2)) / scale)
shapes = [(8, 8), (4, 4), (2, 2)]
off = 0
for l, (h, w) in enumerate(shapes):
    e = np.abs(err[:, off:off + h * w]).max(axis=2).reshape(2, h, w) / scale
    print("level", l, "image 0 per-token max err:\n", np.array2string(e[0], precision=5, suppress_small=True))
    off += h * w
mask, vr, pos = m._masked_constants(d(g["src_mask"]), tuple(shapes), dev, torch.float32)
print("vr", vr.cpu().numpy())
