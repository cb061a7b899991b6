import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import oracle as O
import emrt_b200
from parity import oracle_encdec_pair, l2
dev = torch.device("cuda", 0)
params = O.make_encoder_decoder_params(71, num_enc=4, num_dec=2)
rng = np.random.Generator(np.random.PCG64(72))
tile, B = 256, 2
feats = [torch.from_numpy(O.rng_normal(rng, (B, c, tile // s, tile // s), 0.5)).bfloat16() for c, s in zip((512, 1024, 2048), (8, 16, 32))]
psp = torch.from_numpy(O.rng_normal(rng, (B, 256, 110), 0.5)).bfloat16()
trace = {}
(whs, wmem), (rhs, rmem) = oracle_encdec_pair(params, feats, psp, 4, 2, trace=trace)
m = emrt_b200.EncoderDecoder(110, "sine", False, (512, 1024, 2048), 3, 6, 6, 6, 256, 8, 4, 2, 1024, dropout=0.0)
with torch.no_grad():
    sd = m.state_dict()
    for k in sd: sd[k].copy_(torch.as_tensor(params[k]))
m = m.to(dev)
with torch.no_grad():
    hs_e, mem_e = m([f.to(dev) for f in feats], psp.to(dev))
m.train()
hs_t, mem_t = m([f.to(dev) for f in feats], psp.to(dev))
print("eval : mem", l2(mem_e.float(), wmem), "hs", l2(hs_e.float(), whs))
print("train: mem", l2(mem_t.float(), wmem), "hs", l2(hs_t.float(), whs))
# decoder layers teacher-forced from the oracle trace, both paths
d = lambda t: t.to(torch.bfloat16).to(dev)
shapes = ((32, 32), (16, 16), (8, 8))
c = m._constants(shapes, dev, torch.bfloat16)
t_in, x_in = trace["tgt"], trace["enc"][-1]
for i, layer in enumerate(m.decoder.layers):
    layer.train()
    gt = layer(d(t_in), c["ref_dec"], d(x_in), shapes, None, c["qpos"])
    layer.eval()
    with torch.no_grad():
        ge = layer(d(t_in), c["ref_dec"], d(x_in), shapes, None, c["qpos"])
    print("decoder layer", i, "train vs oracle", l2(gt.float(), trace["dec"][i]), "eval vs oracle", l2(ge.float(), trace["dec"][i]))
    t_in = trace["dec"][i]
