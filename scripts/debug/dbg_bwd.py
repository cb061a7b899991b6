import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import emrt_b200
from emrt_b200 import ops, _lib as L
from oracle import emrt_oracle as O
dev = torch.device("cuda", 0)
tile, B, spread, loc_dtype = 256, 2, 1.0, torch.float16
shapes = [(tile // 8,) * 2, (tile // 16,) * 2, (tile // 32,) * 2]
M, D, P = 8, 32, 6
rng = np.random.Generator(np.random.PCG64(100 + tile + B))
_, Lv = O.level_tables(shapes); Lq = Lv
value = torch.from_numpy(O.rng_normal(rng, (B, Lv, M, D))).bfloat16()
gout = torch.from_numpy(O.rng_normal(rng, (B, Lq, M * D), 3.0)).bfloat16()
bias = O.msda_reset_parameters(M * D, M, 3, P).reshape(1, 1, M, 3, P, 2)
off = torch.from_numpy(bias + O.rng_normal(rng, (B, Lq, M, 3, P, 2), spread)).to(loc_dtype)
attn = torch.from_numpy(rng.uniform(0, 1, size=(B, Lq, M, 3, P)).astype(np.float32)).to(loc_dtype)
ref_t = emrt_b200.get_reference_points(shapes, device=dev)
vd, gd, od, ad = value.to(dev), gout.to(dev), off.to(dev), attn.to(dev)
want = ops.msda_gather_bwd(gd, vd, od, ad, shapes, ref=ref_t, mode=L.LOC_PIXEL_OFFSET)
got = ops.msda_gather_bwd(gd, vd, od, ad, shapes, ref=ref_t, mode=L.LOC_PIXEL_OFFSET | L.QUERY_PIXEL_GRID)
torch.cuda.synchronize()
for i, name in enumerate(["gv", "gl", "ga"]):
    diff = (got[i] - want[i]).abs()
    print(name, "max diff", diff.max().item(), "max want", want[i].abs().max().item(), "n bad", (diff > 1e-3 * want[i].abs().max()).sum().item(), "of", diff.numel())
diff = (got[1] - want[1]).abs()
idx = torch.nonzero(diff > 1e-3 * want[1].abs().max())
ref = ref_t.cpu().numpy()
for r in idx[:24].tolist():
    b, q, m, l, p, c = r
    H, W = shapes[l]
    x = ref[0, q, l, 0] * W - 0.5 + off[b, q, m, l, p, 0].item()
    y = ref[0, q, l, 1] * H - 0.5 + off[b, q, m, l, p, 1].item()
    print(r, "got", got[1][b, q, m, l, p].tolist(), "want", want[1][b, q, m, l, p].tolist(), "x,y", round(float(x), 4), round(float(y), 4), "HW", H, W,
          "ga", got[2][b, q, m, l, p].item(), want[2][b, q, m, l, p].item())
