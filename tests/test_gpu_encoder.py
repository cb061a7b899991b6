"""GPU parity of the encoder-layer glue: LayerNorm fusions, the 3x3 conv branch (SIMT and
tcgen05 implicit GEMM), GroupNorm + GELU + skip, and the whole TransformerEncoderLayer / TransformerEncoder against
the oracle's restatement of transformer_encoder_decoder.py:109-239."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle as O
import emrt_b200
from emrt_b200 import ops, _lib as L

from parity import OWN_TOL, assert_bf16_parity, l2, rounded_params

pytestmark = pytest.mark.gpu


def rel_err(got, want):
    want = torch.as_tensor(want).double()
    return ((got.detach().double().cpu() - want).abs().max() / want.abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("dtype,N", [(torch.float32, 256), (torch.bfloat16, 256), (torch.float32, 1024), (torch.bfloat16, 512)])
def test_residual_layernorm(cuda_dev, dtype, N):
    rng = np.random.Generator(np.random.PCG64(N))
    rows = 777
    x, r, pa = (torch.from_numpy(O.rng_normal(rng, (rows, N))).to(dtype) for _ in range(3))
    g = torch.from_numpy(rng.uniform(0.5, 1.5, size=(N,)).astype(np.float32))
    b = torch.from_numpy(O.rng_normal(rng, (N,), 0.1))
    want = F.layer_norm(x.double() + r.double(), (N,), g.double(), b.double(), 1e-5) + pa.double()
    d = lambda t: t.to(cuda_dev)
    got = ops.residual_layernorm(d(x), d(r), d(g), d(b), post_add=d(pa))
    assert rel_err(got.float(), want) < (1e-5 if dtype == torch.float32 else 1e-2)
    got2 = ops.residual_layernorm(d(x), None, d(g), d(b))
    assert rel_err(got2.float(), F.layer_norm(x.double(), (N,), g.double(), b.double(), 1e-5)) < (1e-5 if dtype == torch.float32 else 1e-2)


def _conv_ref(x_tok, weights, shapes, dtype=torch.float64):
    """per-level F.conv2d on tokens [B, Lv, C] -> tokens (oracle formulation, t_e_d.py:163-196)."""
    B, Lv, C = x_tok.shape
    start, _ = O.level_tables(shapes)
    outs = []
    for l, (h, w) in enumerate(shapes):
        x = x_tok[:, start[l]:start[l] + h * w].to(dtype).permute(0, 2, 1).reshape(B, C, h, w)
        y = F.conv2d(x, weights[l].to(dtype), None, 1, 1)
        outs.append(y.flatten(2).permute(0, 2, 1))
    return torch.cat(outs, 1)


@pytest.mark.parametrize("tile,B", [(512, 3), (256, 3), (256, 4), (128, 5)])
def test_conv3x3_tokens_tcgen05_and_simt(cuda_dev, tile, B):
    """Implicit-GEMM conv (nine shifted TMA loads, zero padding from the tensor map) vs F.conv2d in float64 on the
    bf16-rounded operands; the SIMT kernel on the same data as cross-check.  B odd + a 64-pixel level exercises the
    two-images-per-tile box with an out-of-range image."""
    shapes = [(tile // 8,) * 2, (tile // 16,) * 2, (tile // 32,) * 2]
    C = 256
    rng = np.random.Generator(np.random.PCG64(tile + B))
    _, Lv = O.level_tables(shapes)
    x = torch.from_numpy(O.rng_normal(rng, (B, Lv, C))).bfloat16()
    ws = [torch.from_numpy(O.rng_uniform(rng, (C, C, 3, 3), 0.05)) for _ in shapes]
    want = _conv_ref(x.float(), [w.bfloat16().float() for w in ws], shapes)
    wp = ops.pack_conv3x3_weights([w.to(cuda_dev) for w in ws], torch.bfloat16)
    got_simt = ops.conv3x3_tokens(x.to(cuda_dev), wp, shapes, impl=L.IMPL_SIMT)
    assert rel_err(got_simt.float(), want) < 1e-2
    if tile >= 256:
        got_tc = ops.conv3x3_tokens(x.to(cuda_dev), wp, shapes, impl=L.IMPL_TCGEN05)
        assert rel_err(got_tc.float(), want) < 1e-2
        assert rel_err(got_tc.float(), got_simt.float().cpu()) < 1e-2


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_groupnorm_gelu_residual(cuda_dev, dtype):
    shapes = [(16, 16), (8, 8), (4, 4)]
    B, C = 3, 256
    rng = np.random.Generator(np.random.PCG64(9))
    start, Lv = O.level_tables(shapes)
    conv = torch.from_numpy(O.rng_normal(rng, (B, Lv, C), 2.0) + 0.5).to(dtype)
    x = torch.from_numpy(O.rng_normal(rng, (B, Lv, C))).to(dtype)
    gw = torch.from_numpy(rng.uniform(0.5, 1.5, size=(3, C)).astype(np.float32))
    gb = torch.from_numpy(O.rng_normal(rng, (3, C), 0.1))
    outs = []
    for l, (h, w) in enumerate(shapes):
        c = conv[:, start[l]:start[l] + h * w].double().permute(0, 2, 1).reshape(B, C, h, w)
        y = F.gelu(F.group_norm(c, 32, gw[l].double(), gb[l].double(), 1e-5))
        outs.append(y.flatten(2).permute(0, 2, 1) + x[:, start[l]:start[l] + h * w].double())
    want = torch.cat(outs, 1)
    d = lambda t: t.to(cuda_dev)
    got = ops.groupnorm_gelu_residual(d(conv), d(x), d(gw), d(gb), shapes)
    assert rel_err(got.float(), want) < (1e-4 if dtype == torch.float32 else 1e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_layernorm_with_conv_branch_fused(cuda_dev, dtype):
    """norm2 + the layer's final `src + src_flatten` with the conv branch (GroupNorm + GELU + skip) evaluated inside
    the LayerNorm pass (emrt_groupnorm_stats + emrt_residual_layernorm_gn) against float64 torch."""
    shapes = [(16, 16), (8, 8), (4, 4)]
    B, C = 3, 256
    rng = np.random.Generator(np.random.PCG64(19))
    start, Lv = O.level_tables(shapes)
    mk = lambda std=1.0, shift=0.0: torch.from_numpy(O.rng_normal(rng, (B, Lv, C), std) + np.float32(shift)).to(dtype)
    f, x, conv, src = mk(), mk(), mk(2.0, 0.5), mk()
    lw = torch.from_numpy(rng.uniform(0.5, 1.5, size=(C,)).astype(np.float32))
    lb = torch.from_numpy(O.rng_normal(rng, (C,), 0.1))
    gw = torch.from_numpy(rng.uniform(0.5, 1.5, size=(3, C)).astype(np.float32))
    gb = torch.from_numpy(O.rng_normal(rng, (3, C), 0.1))
    outs = []
    for l, (h, w) in enumerate(shapes):
        c = conv[:, start[l]:start[l] + h * w].double().permute(0, 2, 1).reshape(B, C, h, w)
        y = F.gelu(F.group_norm(c, 32, gw[l].double(), gb[l].double(), 1e-5))
        outs.append(y.flatten(2).permute(0, 2, 1) + src[:, start[l]:start[l] + h * w].double())
    want = F.layer_norm(f.double() + x.double(), (C,), lw.double(), lb.double(), 1e-5) + torch.cat(outs, 1)
    d = lambda t: t.to(cuda_dev)
    st = ops.groupnorm_stats(d(conv), shapes)
    got = ops.residual_layernorm_gn(d(f), d(x), d(lw), d(lb), d(conv), d(src), st, d(gw), d(gb), shapes)
    assert rel_err(got.float(), want) < (1e-4 if dtype == torch.float32 else 1e-2)
    # and it equals the two-kernel formulation
    branch = ops.groupnorm_gelu_residual(d(conv), d(src), d(gw), d(gb), shapes)
    two = ops.residual_layernorm(d(f), d(x), d(lw), d(lb), post_add=branch)
    assert rel_err(got.float(), two.float().cpu()) < (1e-5 if dtype == torch.float32 else 1e-2)


def _load_layer(layer, params, prefix):
    with torch.no_grad():
        sd = layer.state_dict()
        for k in sd:
            sd[k].copy_(torch.as_tensor(params[prefix + k]))
    return layer


def test_encoder_layer_fp32_matches_oracle(cuda_dev):
    shapes = [(16, 16), (8, 8), (4, 4)]
    B, C = 2, 256
    params = O.make_encoder_decoder_params(21, num_enc=1, num_dec=0)
    rng = np.random.Generator(np.random.PCG64(22))
    _, Lv = O.level_tables(shapes)
    src = torch.from_numpy(O.rng_normal(rng, (B, Lv, C)))
    pos = torch.from_numpy(O.rng_normal(rng, (B, Lv, C), 0.5))
    ref = O.encoder_reference_points(shapes, B)
    p64 = {k: torch.as_tensor(v).double() for k, v in params.items()}
    want = O.encoder_layer_forward(p64, "encoder.layers.0.", src.double(), ref.double(), shapes, torch.ones(B, Lv).double(), pos.double())
    layer = _load_layer(emrt_b200.TransformerEncoderLayer(C, 8, 1024, 0.1, "relu", 3, 6), params, "encoder.layers.0.").to(cuda_dev)
    d = lambda t: t.to(cuda_dev)
    got = layer(d(src), d(ref), torch.tensor(shapes), d(torch.ones(B, Lv)), d(pos))
    assert rel_err(got, want) < 2e-4


@pytest.mark.parametrize("tile", [256, 512])
def test_encoder_bf16_two_layers_matches_oracle(cuda_dev, tile):
    """TransformerEncoder (reference points + 2 layers) in bf16 on the B200 path (tcgen05 conv / projections / FFN,
    window-staged gather) vs the float64 oracle on bf16-rounded inputs and weights."""
    shapes = [(tile // 8,) * 2, (tile // 16,) * 2, (tile // 32,) * 2]
    B, C = 2, 256
    params = O.make_encoder_decoder_params(23, num_enc=2, num_dec=0)
    rng = np.random.Generator(np.random.PCG64(24))
    _, Lv = O.level_tables(shapes)
    r16 = lambda a: torch.as_tensor(a).bfloat16()
    src = r16(O.rng_normal(rng, (B, Lv, C), 0.5))
    pos = r16(O.rng_normal(rng, (1, Lv, C), 0.5))
    p64 = rounded_params(params)
    ref = O.encoder_reference_points(shapes, B).double()
    def run(keep=None):
        x = src.double()
        for i in range(2):
            x = O.encoder_layer_forward(p64, f"encoder.layers.{i}.", x, ref, shapes, torch.ones(B, Lv).double(),
                                        pos.double().expand(B, -1, -1))
            if keep is not None:
                keep.append(x)
        return x
    want = run()
    per_layer = []
    with O.kernel_storage_rounding():
        rounded = run(per_layer)
    layer = emrt_b200.TransformerEncoderLayer(C, 8, 1024, 0.1, "relu", 3, 6)
    enc = emrt_b200.TransformerEncoder(layer, 2)
    for i in range(2):
        _load_layer(enc.layers[i], params, f"encoder.layers.{i}.")
    enc = enc.to(cuda_dev)
    got = enc(src.to(cuda_dev), torch.tensor(shapes), None, pos.to(cuda_dev))
    assert got.dtype == torch.bfloat16
    # tests/parity.py: each layer on the same-rounding oracle's own input within 1e-3 (the kernels' own error); end to end no
    # further from the exact run than the storage formats alone are
    ref_d = emrt_b200.get_reference_points(shapes, device=cuda_dev)
    x_in = src.double()
    for i in range(2):
        y = enc.layers[i](x_in.bfloat16().to(cuda_dev), ref_d, shapes, None, pos.to(cuda_dev))
        own = l2(y.float(), per_layer[i])
        print(f"layer {i}: kernels vs same-rounding oracle on its input {own:.1e}")
        assert own < OWN_TOL
        x_in = per_layer[i]
    assert_bf16_parity(got.float(), want, rounded, f"2-layer encoder, tile {tile}")


@pytest.mark.parametrize("B,shapes", [(3, [(32, 32), (16, 16), (8, 8)]), (5, [(16, 32), (8, 16), (4, 8)]), (2, [(64, 64), (32, 32), (16, 16)])])
def test_conv_epilogue_groupnorm_statistics(cuda_dev, B, shapes):
    """emrt_conv3x3_tokens_stats_fwd: the same conv output bit for bit, and GroupNorm sums that equal the separate statistics
    kernel's up to fp32 summation order (both sum the stored bf16 values) — single-CTA and CTA-pair forms."""
    import os
    rng = np.random.Generator(np.random.PCG64(B))
    Lv = sum(h * w for h, w in shapes)
    x = torch.from_numpy(O.rng_normal(rng, (B, Lv, 256))).to(cuda_dev).bfloat16()
    ws = [torch.from_numpy(O.rng_uniform(rng, (256, 256, 3, 3), 0.05)).to(cuda_dev) for _ in shapes]
    wp = ops.pack_conv3x3_weights(ws, torch.bfloat16)
    want_y = ops.conv3x3_tokens(x, wp, shapes)
    want_s = ops.groupnorm_stats(want_y, shapes, groups=32)[: 2 * B * len(shapes) * 32]
    for one_cta in (False, True):
        if one_cta:
            os.environ["EMRT_CONV_1CTA"] = "1"
        try:
            y, st = ops.conv3x3_tokens_stats(x, wp, shapes)
            y2, st2 = ops.conv3x3_tokens_stats(x, wp, shapes)
        finally:
            os.environ.pop("EMRT_CONV_1CTA", None)
        assert torch.equal(y, want_y)
        got_s = st[: 2 * B * len(shapes) * 32]
        assert torch.equal(got_s, st2[: 2 * B * len(shapes) * 32])          # deterministic
        scale = want_s.abs().max().item()
        assert (got_s - want_s).abs().max().item() <= 2e-5 * scale


@pytest.mark.parametrize("tile,B,ctas", [(256, 2, 4), (512, 3, 64), (256, 5, 2)])
def test_encoder_layer_conv_beside_gather_is_bit_equal(cuda_dev, tile, B, ctas):
    """TransformerEncoderLayer.overlap_conv: the 3x3 conv on `ctas` SMs of a side stream, started by the event the fused MSDA
    call records right before its gather — against the same layer in order on one stream (overlap_conv = 0).  Same kernels,
    persistent tile loops, per-tile statistics partials: the outputs must be identical bit for bit, call after call (the
    events order the two streams: a race would show as a difference)."""
    shapes = [(tile // 8,) * 2, (tile // 16,) * 2, (tile // 32,) * 2]
    C = 256
    params = O.make_encoder_decoder_params(29, num_enc=2, num_dec=0)
    rng = np.random.Generator(np.random.PCG64(30))
    _, Lv = O.level_tables(shapes)
    src = torch.as_tensor(O.rng_normal(rng, (B, Lv, C), 0.5)).bfloat16().to(cuda_dev)
    pos = torch.as_tensor(O.rng_normal(rng, (1, Lv, C), 0.5)).bfloat16().to(cuda_dev)
    layer = emrt_b200.TransformerEncoderLayer(C, 8, 1024, 0.1, "relu", 3, 6)
    enc = emrt_b200.TransformerEncoder(layer, 2)
    for i in range(2):
        _load_layer(enc.layers[i], params, f"encoder.layers.{i}.")
    enc = enc.to(cuda_dev)
    assert enc.layers[0].overlap_conv > 0
    outs = {}
    for k in (0, ctas, 0, ctas):
        for lyr in enc.layers:
            lyr.overlap_conv = k
        outs.setdefault(k, []).append(enc(src, torch.tensor(shapes), None, pos))
    torch.cuda.synchronize()
    for o in outs[ctas] + outs[0][1:]:
        assert torch.equal(o, outs[0][0])
    assert torch.isfinite(outs[0][0].float()).all()


def test_conv3x3_stats_on_a_share_of_the_sms(cuda_dev):
    """emrt_conv3x3_tokens_stats_part_fwd: any max_ctas (odd values round down to whole CTA pairs) gives the conv output and the
    GroupNorm sums of the full-grid launch, bit for bit."""
    shapes = [(64, 64), (32, 32), (16, 16)]
    B, C = 4, 256
    g = torch.Generator(device="cpu").manual_seed(31)
    x = (torch.randn((B, sum(h * w for h, w in shapes), C), generator=g) * 0.5).bfloat16().to(cuda_dev)
    w = ops.pack_conv3x3_weights([(torch.randn((C, C, 3, 3), generator=g) * 0.02).to(cuda_dev) for _ in shapes], torch.bfloat16)
    y0, s0 = ops.conv3x3_tokens_stats(x, w, shapes)
    n = 2 * B * len(shapes) * 32
    for k in (1, 2, 7, 40, 1000):
        y, s = ops.conv3x3_tokens_stats(x, w, shapes, max_ctas=k)
        assert torch.equal(y, y0) and torch.equal(s[:n], s0[:n]), k
