"""Generates tests/golden/*.npz from the CPU oracle (oracle/emrt_oracle.py).

These two fixtures are regression pins of the oracle itself (float64-evaluated where noted).  The vectors that come
from the reference's own code are the ref_*.npz files written by make_reference_vectors.py.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def msda_case():
    rng = np.random.Generator(np.random.PCG64(7))
    shapes = [(8, 6), (4, 3), (2, 2)]                  # non-square levels
    B, C, M, P = 2, 64, 2, 6
    _, Lv = O.level_tables(shapes)
    Lq = 37                                             # Lq != Lv (decoder-style)
    params = O.make_msda_params(99, C, M, len(shapes), P, offset_std=0.15)
    q = O.rng_normal(rng, (B, Lq, C))
    v = O.rng_normal(rng, (B, Lv, C))
    ref = rng.uniform(0, 1, size=(B, Lq, len(shapes), 2)).astype(np.float32)
    mask = (rng.uniform(size=(B, Lv)) > 0.1).astype(np.float32)
    vp, loc, aw = O.msda_intermediates(params, q, ref, v, shapes, mask, M, P, dtype=torch.float64)
    core = O.gather_corner_loop(vp.numpy(), shapes, loc.numpy(), aw.numpy())
    out = O.msda_forward(params, q, ref, v, shapes, mask, M, P, dtype=torch.float64)
    np.savez_compressed(os.path.join(OUT, "msda_small.npz"), shapes=np.array(shapes), query=q, value=v, ref=ref,
                        mask=mask, value_proj=vp.numpy().astype(np.float32), loc=loc.numpy().astype(np.float32),
                        attn=aw.numpy().astype(np.float32), core=core.astype(np.float32),
                        out=out.numpy().astype(np.float32), **{"p." + k: a for k, a in params.items()})


def slide_case():
    rng = np.random.Generator(np.random.PCG64(11))
    nc, crop, stride = 6, (32, 24), (20, 16)            # (w, h) order like the reference
    imgs = [O.rng_normal(rng, (3, 50, 70)), O.rng_normal(rng, (3, 41, 33))]
    wmat = O.rng_normal(rng, (nc, 3), 0.7)

    def model(batch):                                   # a fixed per-pixel linear "model": logits = W @ rgb
        return (torch.einsum("oc,nchw->nohw", torch.from_numpy(wmat), batch),)
    logits = O.slide_inference(model, [torch.from_numpy(i) for i in imgs], crop, stride, nc)
    ori = [(50, 70), (60, 50)]
    preds = [O.ss_inference_tail(l, ori[i]).numpy() for i, l in enumerate(logits)]
    np.savez_compressed(os.path.join(OUT, "slide_small.npz"), img0=imgs[0], img1=imgs[1], wmat=wmat,
                        crop=np.array(crop), stride=np.array(stride), logit0=logits[0].numpy(), logit1=logits[1].numpy(),
                        pred0=preds[0], pred1=preds[1], ori=np.array(ori))


if __name__ == "__main__":
    msda_case()
    slide_case()
    print("wrote", sorted(f for f in os.listdir(OUT) if f.endswith(".npz")))
