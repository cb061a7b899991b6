"""GPU parity of emrt_ffn_fused_fwd: the encoder layer's forward_ffn + norm2 (+ conv branch and the layer's final add) as one
tcgen05 kernel whose hidden activations stay on the SM (transformer_encoder_decoder.py:157-160,187-189,203)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle as O
from emrt_b200 import ops, _lib as L

pytestmark = pytest.mark.gpu


def rel_err(got, want):
    want = torch.as_tensor(want).double().cpu()
    return ((got.detach().double().cpu() - want).abs().max() / want.abs().max().clamp_min(1e-30)).item()


def l2_err(got, want):
    want = torch.as_tensor(want).double().cpu()
    return ((got.detach().double().cpu() - want).norm() / want.norm().clamp_min(1e-30)).item()


def bf16_round(t):
    return t.float().bfloat16().double()


def make_case(rng, dev, B, shapes, d_ff, gn):
    C = 256
    Lv = sum(h * w for h, w in shapes)
    rows = B * Lv
    d = lambda a: torch.from_numpy(a).to(dev)
    x = d(O.rng_normal(rng, (B, Lv, C))).bfloat16()
    w1 = d(O.rng_uniform(rng, (C, d_ff), (6.0 / (C + d_ff)) ** 0.5))
    w2 = d(O.rng_uniform(rng, (d_ff, C), (6.0 / (C + d_ff)) ** 0.5))
    b1 = d(O.rng_uniform(rng, (d_ff,), 0.2))
    b2 = d(O.rng_uniform(rng, (C,), 0.2))
    g = d(rng.uniform(0.5, 1.5, size=(C,)).astype(np.float32))
    bt = d(O.rng_normal(rng, (C,), 0.1))
    w1p = torch.empty((d_ff, C), dtype=torch.bfloat16, device=dev)
    w2p = torch.empty((C, d_ff), dtype=torch.bfloat16, device=dev)
    ops.pack_weight(w1, w1p)
    ops.pack_weight(w2, w2p)
    case = dict(x=x, w1p=w1p, w2p=w2p, b1=b1, b2=b2, g=g, bt=bt, rows=rows, shapes=shapes, gn=None)
    if gn:
        conv = d(O.rng_normal(rng, (B, Lv, C))).bfloat16()
        skip = d(O.rng_normal(rng, (B, Lv, C))).bfloat16()
        gw = d(rng.uniform(0.5, 1.5, size=(len(shapes), C)).astype(np.float32))
        gb = d(O.rng_normal(rng, (len(shapes), C), 0.1))
        stats = ops.groupnorm_stats(conv, shapes, groups=32)
        case["gn"] = dict(conv=conv, skip=skip, stats=stats, gamma=gw, beta=gb, shapes=shapes, groups=32, eps=1e-5)
    return case


def reference(case):
    """float64 on the bf16 operands with the kernel's two roundings: the hidden chunk and the pre-LayerNorm sum"""
    x = case["x"].double()
    h = bf16_round(torch.relu(x @ case["w1p"].double().T + case["b1"].double()))
    t = bf16_round(h @ case["w2p"].double().T + case["b2"].double() + x)
    y = F.layer_norm(t, (256,), case["g"].double(), case["bt"].double(), 1e-5)
    gn = case["gn"]
    if gn is not None:
        B, Lv, C = x.shape
        conv = gn["conv"].double()
        start = 0
        br = torch.empty_like(conv)
        for l, (hh, ww) in enumerate(case["shapes"]):
            n = hh * ww
            cl = conv[:, start:start + n].reshape(B, n, 32, C // 32)
            mean = cl.mean(dim=(1, 3), keepdim=True)
            var = cl.var(dim=(1, 3), unbiased=False, keepdim=True)
            nl = ((cl - mean) / torch.sqrt(var + 1e-5)).reshape(B, n, C) * gn["gamma"][l].double() + gn["beta"][l].double()
            br[:, start:start + n] = F.gelu(nl)
            start += n
        y = y + br + gn["skip"].double()
    return y


@pytest.mark.parametrize("B,shapes,d_ff,gn", [
    (3, [(16, 16), (8, 8), (4, 4)], 1024, True),       # 1008 rows: 7 full tiles + a partial one
    (2, [(16, 16), (8, 8), (4, 4)], 256, True),        # two hidden chunks
    (1, [(8, 8), (4, 4), (2, 2)], 128, False),         # one chunk, less than a tile, plain LayerNorm
    (5, [(8, 16), (4, 8), (2, 4)], 384, False),         # three chunks: the hidden k-block ring (3 slots) wraps differently per tile
    (60, [(16, 16), (8, 8), (4, 4)], 1024, True),      # 20160 rows: more tiles than SMs (two tiles on some CTAs)
])
def test_ffn_fused_matches_float64_and_two_kernel_form(cuda_dev, B, shapes, d_ff, gn):
    rng = np.random.Generator(np.random.PCG64(B * 1000 + d_ff))
    case = make_case(rng, cuda_dev, B, shapes, d_ff, gn)
    got = ops.ffn_fused(case["x"], case["w1p"], case["b1"], case["w2p"], case["b2"], case["g"], case["bt"], gn_branch=case["gn"])
    assert got.dtype == torch.bfloat16 and got.shape == case["x"].shape
    want = reference(case)
    # the output's own bf16 rounding (2^-9) plus flips of the two inner roundings where fp32 and float64 accumulation
    # disagree by one ulp
    assert rel_err(got.float(), want) < 8e-3 and l2_err(got.float(), want) < 3e-3
    # the two-kernel composition (hidden tensor through HBM, LayerNorm on the fp32 accumulator)
    h = ops.linear(case["x"], case["w1p"], case["b1"], w_transposed=True, epilogue=L.EPI_RELU)
    two = ops.linear(h, case["w2p"], case["b2"], w_transposed=True, epilogue=L.EPI_RESIDUAL_LN, residual=case["x"],
                     ln_gamma=case["g"], ln_beta=case["bt"], gn_branch=case["gn"]) if (gn and d_ff > 256) else None
    if two is not None:
        assert l2_err(got.float(), two.float()) < 6e-3
    # deterministic
    again = ops.ffn_fused(case["x"], case["w1p"], case["b1"], case["w2p"], case["b2"], case["g"], case["bt"], gn_branch=case["gn"])
    assert torch.equal(got, again)


def test_ffn_fused_rejects_what_it_is_not_built_for(cuda_dev):
    x = torch.zeros((4, 128), dtype=torch.bfloat16, device=cuda_dev)
    w1 = torch.zeros((64, 128), dtype=torch.bfloat16, device=cuda_dev)
    w2 = torch.zeros((128, 64), dtype=torch.bfloat16, device=cuda_dev)
    v = torch.zeros((128,), dtype=torch.float32, device=cuda_dev)
    with pytest.raises(L.EmrtError):
        ops.ffn_fused(x, w1, v[:64].contiguous(), w2, v, v, v)


def test_ffn_fused_single_cta_form_equals_cta_pair_form(cuda_dev, monkeypatch):
    """EMRT_FFN_1CTA=1 (cta_group::1, one CTA per row tile) and the default CTA-pair form (cta_group::2) run the same
    arithmetic in the same order per output element: bit-equal."""
    rng = np.random.Generator(np.random.PCG64(77))
    case = make_case(rng, cuda_dev, 7, [(16, 16), (8, 8), (4, 4)], 1024, True)      # 2352 rows: an odd number of tiles (19)
    args = (case["x"], case["w1p"], case["b1"], case["w2p"], case["b2"], case["g"], case["bt"])
    pair = ops.ffn_fused(*args, gn_branch=case["gn"])
    monkeypatch.setenv("EMRT_FFN_1CTA", "1")
    single = ops.ffn_fused(*args, gn_branch=case["gn"])
    monkeypatch.delenv("EMRT_FFN_1CTA")
    assert torch.equal(pair, single)
    assert l2_err(pair.float(), reference(case)) < 3e-3
