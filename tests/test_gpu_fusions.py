"""GPU parity of the round-2 fusions and autograd fixes:
  * EMRT_EPI_RESIDUAL_LN — Linear + residual + LayerNorm in one tcgen05 kernel (t_e_d.py:106,199-200,157-160)
  * x2 — with_pos_embed folded into a projection as extra k-blocks of the same accumulation (:154-155,198)
  * the reference-point gradient (:98-102,466-467), the differentiable core function (utils.py:64-97)
  * bf16 MSDA with a non-EMRT point configuration (4 levels x 4 points), geometry-based kernel selection."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle as O
import emrt_b200
from emrt_b200 import ops, _lib as L
from emrt_b200.msda import is_pixel_grid

pytestmark = pytest.mark.gpu


def rel_err(got, want):
    want = torch.as_tensor(want).double()
    return ((got.detach().double().cpu() - want).abs().max() / want.abs().max().clamp_min(1e-30)).item()


def l2_err(got, want):
    want = torch.as_tensor(want).double()
    return ((got.detach().double().cpu() - want).norm() / want.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("rows,K", [(777, 256), (128 * 150 + 5, 256), (4096, 1024), (33, 1024), (110 * 3, 256)])
def test_linear_residual_layernorm_fused(cuda_dev, rows, K):
    """y = LN(x W + b + res) from the fp32 accumulator vs float64 on the bf16-rounded operands: the only rounding left is
    the output's (2^-9 relative), so the bound is tighter than the two-kernel form's."""
    N = 256
    rng = np.random.Generator(np.random.PCG64(rows + K))
    x = torch.from_numpy(O.rng_normal(rng, (rows, K))).bfloat16()
    res = torch.from_numpy(O.rng_normal(rng, (rows, N))).bfloat16()
    w = torch.from_numpy(O.rng_uniform(rng, (K, N), (6.0 / (K + N)) ** 0.5))
    b = torch.from_numpy(O.rng_uniform(rng, (N,), 0.1))
    g = torch.from_numpy(rng.uniform(0.5, 1.5, size=(N,)).astype(np.float32))
    bt = torch.from_numpy(O.rng_normal(rng, (N,), 0.1))
    d = lambda t: t.to(cuda_dev)
    wp = torch.empty((N, K), dtype=torch.bfloat16, device=cuda_dev)
    ops.pack_weight(d(w), wp)
    want = F.layer_norm(x.double() @ w.bfloat16().double() + b.double() + res.double(), (N,), g.double(), bt.double(), 1e-5)
    got = ops.linear(d(x), wp, d(b), w_transposed=True, epilogue=L.EPI_RESIDUAL_LN, residual=d(res), ln_gamma=d(g),
                     ln_beta=d(bt))
    assert got.dtype == torch.bfloat16 and tuple(got.shape) == (rows, N)
    assert rel_err(got.float(), want) < 6e-3 and l2_err(got.float(), want) < 3e-3
    # against the two-kernel composition (which rounds the projection to bf16 first)
    y = ops.linear(d(x), wp, d(b), w_transposed=True)
    two = ops.residual_layernorm(y, d(res), d(g), d(bt))
    assert l2_err(got.float(), two.float().cpu()) < 6e-3
    # in place over the residual and over x's buffer when shapes allow
    r2 = d(res).clone()
    got2 = ops.linear(d(x), wp, d(b), w_transposed=True, epilogue=L.EPI_RESIDUAL_LN, residual=r2, ln_gamma=d(g),
                      ln_beta=d(bt), out=r2)
    assert torch.equal(got2, got)
    if K == N:
        x2 = d(x).clone()
        got3 = ops.linear(x2, wp, d(b), w_transposed=True, epilogue=L.EPI_RESIDUAL_LN, residual=d(res), ln_gamma=d(g),
                          ln_beta=d(bt), out=x2)
        assert torch.equal(got3, got)


def test_linear_residual_layernorm_rejects_unsupported(cuda_dev):
    x = torch.zeros((8, 256), dtype=torch.bfloat16, device=cuda_dev)
    w = torch.zeros((128, 256), dtype=torch.bfloat16, device=cuda_dev)
    g = torch.ones((128,), device=cuda_dev)
    with pytest.raises(L.EmrtError, match="N = 256"):
        ops.linear(x, w, None, w_transposed=True, epilogue=L.EPI_RESIDUAL_LN,
                   residual=torch.zeros((8, 128), dtype=torch.bfloat16, device=cuda_dev), ln_gamma=g, ln_beta=g)
    xf = torch.zeros((8, 256), device=cuda_dev)
    wf = torch.zeros((256, 256), device=cuda_dev)
    with pytest.raises(L.EmrtError):          # the fp32 parity path composes the two kernels; the SIMT GEMM has no LN epilogue
        ops.linear(xf, wf, None, epilogue=L.EPI_RESIDUAL_LN, residual=xf, ln_gamma=g, ln_beta=g)


@pytest.mark.parametrize("B,period,N,epi", [(3, 1344, 256, "none"), (2, 5376, 432, "qproj"), (5, 110, 512, "none"),
                                            (4, 110, 432, "qproj"), (3, 1344, 432, "qproj")])
def test_linear_x2_is_with_pos_embed(cuda_dev, B, period, N, epi):
    """(x + pos) W + b evaluated as x W + pos W in one accumulator vs float64 on the bf16 operands; periods that are not a
    multiple of the 128-row tile (1344, 110) exercise the cyclic continuation."""
    K = 256
    rng = np.random.Generator(np.random.PCG64(B * period + N))
    x = torch.from_numpy(O.rng_normal(rng, (B, period, K))).bfloat16()
    pos = torch.from_numpy(O.rng_normal(rng, (1, period, K))).bfloat16()
    w = torch.from_numpy(O.rng_uniform(rng, (K, N), 0.06))
    b = torch.from_numpy(O.rng_uniform(rng, (N,), 0.1))
    d = lambda t: t.to(cuda_dev)
    wp = torch.empty((N, K), dtype=torch.bfloat16, device=cuda_dev)
    ops.pack_weight(d(w), wp)
    pos_d = d(pos)
    cyc = ops.cyclic_rows_cached(pos_d)
    assert ops.cyclic_rows_cached(pos_d) is cyc and tuple(cyc.shape) == (period + 127, K)
    raw = (x.double() + pos.double()) @ w.bfloat16().double() + b.double()
    if epi == "none":
        got = ops.linear(d(x), wp, d(b), w_transposed=True, y_dtype=torch.float32, x2=cyc, x2_period=period)
        assert rel_err(got, raw) < 2e-5
        added = ops.add_bcast(d(x), pos_d)                       # the un-fused form rounds x + pos to bf16
        old = ops.linear(added, wp, d(b), w_transposed=True, y_dtype=torch.float32)
        assert rel_err(old, raw) < 1e-2 and rel_err(got, raw) < rel_err(old, raw)
    else:
        off, attn = ops.linear(d(x), wp, d(b), w_transposed=True, y_dtype=torch.float16, epilogue=L.EPI_MSDA_QPROJ,
                               qproj_group=18, x2=cyc, x2_period=period)
        tp = N // 3
        assert rel_err(off.float(), raw[..., :2 * tp]) < 2e-3
        want_attn = torch.softmax(raw[..., 2 * tp:].reshape(B, period, tp // 18, 18), -1).reshape(B, period, tp)
        assert rel_err(attn.float(), want_attn) < 2e-3
        # the same addend as a precomputed row bias (pos W + b, fp16, cyclic) added in the epilogue: no extra MMA work
        tab = ops.linear(pos_d.reshape(period, K).float(), d(w.bfloat16().float()), d(b), y_dtype=torch.float16, impl=L.IMPL_SIMT)
        tab = ops.cyclic_rows(tab, dtype=torch.float16)
        off2, attn2 = ops.linear(d(x), wp, None, w_transposed=True, y_dtype=torch.float16, epilogue=L.EPI_MSDA_QPROJ,
                                 qproj_group=18, row_bias=tab, row_bias_period=period)
        assert rel_err(off2.float(), raw[..., :2 * tp]) < 2e-3 and rel_err(attn2.float(), want_attn) < 2e-3


def test_reference_point_gradient_matches_oracle(cuda_dev):
    """d loss / d reference_points through MSDeformableAttention (t_e_d.py:98-102): batch-shared (the decoder's trained
    sigmoid(Linear(query_pos_embed)), :466-467) and per-batch, fp32 path vs torch autograd through the oracle."""
    shapes = [(12, 10), (6, 5), (3, 3)]
    B, C, M, P, Lq = 3, 64, 4, 2, 17
    _, Lv = O.level_tables(shapes)
    rng = np.random.Generator(np.random.PCG64(5))
    params = O.make_msda_params(77, C, M, len(shapes), P)
    q = torch.from_numpy(O.rng_normal(rng, (B, Lq, C)))
    v = torch.from_numpy(O.rng_normal(rng, (B, Lv, C)))
    g_out = torch.from_numpy(O.rng_normal(rng, (B, Lq, C)))
    for ref_b in (1, B):
        ref = torch.from_numpy(rng.uniform(0.1, 0.9, size=(ref_b, Lq, len(shapes), 2)).astype(np.float32))
        # oracle autograd
        ref_o = ref.clone().double().requires_grad_(True)
        po = {k: torch.from_numpy(a).double() for k, a in params.items()}
        out_o = O.msda_forward(po, q.double(), ref_o.expand(B, -1, -1, -1), v.double(), shapes, None, M, P, torch.float64)
        (out_o * g_out.double()).sum().backward()
        # CUDA path
        attn = emrt_b200.MSDeformableAttention(C, M, len(shapes), P).to(cuda_dev)
        with torch.no_grad():
            for name, arr in params.items():
                mod, leaf = name.split(".")
                getattr(getattr(attn, mod), leaf).copy_(torch.from_numpy(arr))
        ref_c = ref.to(cuda_dev).requires_grad_(True)
        q_c = q.to(cuda_dev).requires_grad_(True)
        out_c = attn(q_c, ref_c, v.to(cuda_dev), shapes)
        assert rel_err(out_c, out_o.detach()) < 1e-4
        (out_c * g_out.to(cuda_dev)).sum().backward()
        assert ref_c.grad is not None and tuple(ref_c.grad.shape) == tuple(ref.shape)
        assert rel_err(ref_c.grad, ref_o.grad) < 2e-4, f"ref batch {ref_b}"


def test_reference_point_gradient_bf16_pixel_mode(cuda_dev):
    """Same in the bf16 path (PIXEL_OFFSET mode: d x / d ref_x = W_l)."""
    shapes = [(16, 16), (8, 8), (4, 4)]
    B, C, M, P, Lq = 2, 256, 8, 6, 110
    _, Lv = O.level_tables(shapes)
    rng = np.random.Generator(np.random.PCG64(6))
    params = O.make_msda_params(78, C, M, len(shapes), P)
    q = torch.from_numpy(O.rng_normal(rng, (B, Lq, C))).bfloat16()
    v = torch.from_numpy(O.rng_normal(rng, (B, Lv, C))).bfloat16()
    g_out = torch.from_numpy(O.rng_normal(rng, (B, Lq, C)))
    ref = torch.from_numpy(rng.uniform(0.2, 0.8, size=(1, Lq, len(shapes), 2)).astype(np.float32))
    ref_o = ref.clone().double().requires_grad_(True)
    po = {k: torch.from_numpy(a).bfloat16().double() if a.ndim == 2 else torch.from_numpy(a).double() for k, a in params.items()}
    out_o = O.msda_forward(po, q.double(), ref_o.expand(B, -1, -1, -1), v.double(), shapes, None, M, P, torch.float64)
    (out_o * g_out.double()).sum().backward()
    attn = emrt_b200.MSDeformableAttention(C, M, len(shapes), P).to(cuda_dev)
    with torch.no_grad():
        for name, arr in params.items():
            mod, leaf = name.split(".")
            getattr(getattr(attn, mod), leaf).copy_(torch.from_numpy(arr))
    ref_c = ref.to(cuda_dev).requires_grad_(True)
    out_c = attn(q.to(cuda_dev).requires_grad_(True), ref_c, v.to(cuda_dev), shapes)
    (out_c.float() * g_out.to(cuda_dev)).sum().backward()
    # sums of signed per-point terms over 8 heads x 6 points x 2 images: bf16 value differences and fp16 weights leave
    # ~4 % of the (partly cancelling) total
    assert l2_err(ref_c.grad, ref_o.grad) < 8e-2


def test_core_func_is_differentiable(cuda_dev):
    """deformable_attention_core_func carries a grad_fn like the reference's composition of Paddle ops (utils.py:64-97)."""
    shapes = [(9, 7), (5, 4)]
    B, M, D, Lq, P = 2, 3, 16, 11, 4
    _, Lv = O.level_tables(shapes)
    rng = np.random.Generator(np.random.PCG64(8))
    value = torch.from_numpy(O.rng_normal(rng, (B, Lv, M, D)))
    loc = torch.from_numpy(rng.uniform(-0.1, 1.1, size=(B, Lq, M, len(shapes), P, 2)).astype(np.float32))
    aw = torch.softmax(torch.from_numpy(O.rng_normal(rng, (B, Lq, M, len(shapes) * P))), -1).reshape(B, Lq, M, len(shapes), P)
    g_out = torch.from_numpy(O.rng_normal(rng, (B, Lq, M * D)))
    vo, lo, ao = (t.clone().double().requires_grad_(True) for t in (value, loc, aw))
    out_o = O.deformable_attention_core_func(vo, shapes, lo, ao)
    (out_o * g_out.double()).sum().backward()
    vc, lc, ac = (t.to(cuda_dev).requires_grad_(True) for t in (value, loc, aw))
    out_c = emrt_b200.deformable_attention_core_func(vc, torch.tensor(shapes), lc, ac)
    assert out_c.grad_fn is not None
    (out_c * g_out.to(cuda_dev)).sum().backward()
    assert rel_err(out_c, out_o.detach()) < 1e-5
    for got, want in ((vc.grad, vo.grad), (lc.grad, lo.grad), (ac.grad, ao.grad)):
        assert rel_err(got, want) < 2e-4
    with torch.no_grad():
        assert emrt_b200.deformable_attention_core_func(vc, shapes, lc, ac).grad_fn is None


@pytest.mark.parametrize("nL,P", [(4, 4), (2, 3)])
def test_msda_bf16_other_point_configurations(cuda_dev, nL, P):
    """gemm_impl AUTO with a configuration the fused MSDA_QPROJ epilogue is not built for — the constructor's own default
    (4 levels x 4 points, also the reference's, t_e_d.py:22) — runs the generic tcgen05 GEMM + emrt_msda_softmax_loc."""
    shapes = [(16, 16), (8, 8), (4, 4), (2, 2)][:nL]
    B, C, M, Lq = 2, 256, 8, 37
    _, Lv = O.level_tables(shapes)
    rng = np.random.Generator(np.random.PCG64(nL * 10 + P))
    params = O.make_msda_params(5, C, M, nL, P)
    q = torch.from_numpy(O.rng_normal(rng, (B, Lq, C)))
    v = torch.from_numpy(O.rng_normal(rng, (B, Lv, C)))
    ref = torch.from_numpy(rng.uniform(0.1, 0.9, size=(B, Lq, nL, 2)).astype(np.float32))
    p64 = {k: (torch.from_numpy(a).bfloat16().double() if a.ndim == 2 else torch.from_numpy(a).double()) for k, a in params.items()}
    want = O.msda_forward(p64, q.bfloat16().double(), ref.double(), v.bfloat16().double(), shapes, None, M, P, torch.float64)
    attn = emrt_b200.MSDeformableAttention(C, M, nL, P).to(cuda_dev)
    assert not attn.fused_qproj_ok()
    with torch.no_grad():
        for name, arr in params.items():
            mod, leaf = name.split(".")
            getattr(getattr(attn, mod), leaf).copy_(torch.from_numpy(arr))
        got = attn(q.to(cuda_dev).bfloat16(), ref.to(cuda_dev), v.to(cuda_dev).bfloat16(), shapes)
    assert l2_err(got.float(), want) < 1e-2
    # and through autograd
    qg = q.to(cuda_dev).bfloat16().requires_grad_(True)
    out = attn(qg, ref.to(cuda_dev), v.to(cuda_dev).bfloat16(), shapes)
    out.float().sum().backward()
    assert qg.grad is not None and torch.isfinite(qg.grad.float()).all()
    assert l2_err(out.float(), want) < 1e-2


def test_pixel_grid_is_detected_from_the_tensor(cuda_dev):
    """The window-staged gather is chosen from the reference points' VALUES (the reference's own get_reference_points
    output, a clone, a .to() copy), not from an attribute of the tensor object."""
    shapes = [(16, 16), (8, 8), (4, 4)]
    _, Lv = O.level_tables(shapes)
    tagged = emrt_b200.refpoints.get_reference_points(shapes, device=cuda_dev)
    plain = O.encoder_reference_points(shapes, 1).to(cuda_dev)         # the oracle's restatement of t_e_d.py:213-228
    assert not hasattr(plain, "pixel_grid")
    assert is_pixel_grid(tagged, shapes, Lv, Lv) and is_pixel_grid(plain, shapes, Lv, Lv)
    assert is_pixel_grid(plain.clone(), shapes, Lv, Lv) and is_pixel_grid(plain.expand(3, -1, -1, -1).contiguous(), shapes, Lv, Lv)
    rnd = torch.rand_like(plain)
    assert not is_pixel_grid(rnd, shapes, Lv, Lv)
    assert not is_pixel_grid(plain[:, :110].contiguous(), shapes, 110, Lv)
    # same results either way (the flag is a locality hint): MSDA with the plain tensor == with the tagged one, bit for bit
    rng = np.random.Generator(np.random.PCG64(3))
    attn = emrt_b200.MSDeformableAttention(256, 8, 3, 6).to(cuda_dev)
    params = O.make_msda_params(11, 256, 8, 3, 6)
    with torch.no_grad():
        for name, arr in params.items():
            mod, leaf = name.split(".")
            getattr(getattr(attn, mod), leaf).copy_(torch.from_numpy(arr))
        x = torch.from_numpy(O.rng_normal(rng, (2, Lv, 256))).to(cuda_dev).bfloat16()
        a = attn(x, tagged, x, shapes)
        ops.reset_launch_count()
        b = attn(x, plain, x, shapes)
    assert torch.equal(a, b)


def test_encoder_layer_fused_equals_unfused(cuda_dev):
    """TransformerEncoderLayer with the round-2 fusions (pos folded into the query projection, norm1 in the output
    projection's epilogue) against the same layer composed of the separate kernels, and both against the oracle."""
    shapes = [(32, 32), (16, 16), (8, 8)]
    B, C = 2, 256
    _, Lv = O.level_tables(shapes)
    rng = np.random.Generator(np.random.PCG64(21))
    params = O.make_encoder_decoder_params(31, num_enc=1, num_dec=0)
    layer = emrt_b200.TransformerEncoderLayer(C, 8, 1024, 0.1, "relu", 3, 6)
    with torch.no_grad():
        sd = layer.state_dict()
        for k in sd:
            sd[k].copy_(torch.as_tensor(params["encoder.layers.0." + k]))
    layer = layer.to(cuda_dev)
    src = torch.from_numpy(O.rng_normal(rng, (B, Lv, C), 0.5))
    pos = torch.from_numpy(O.rng_normal(rng, (1, Lv, C), 0.5))
    ref = emrt_b200.refpoints.get_reference_points(shapes, device=cuda_dev)
    r16 = lambda a: torch.as_tensor(a).bfloat16()
    p64 = {k: (r16(v).double() if v.ndim >= 2 else torch.as_tensor(v).double()) for k, v in params.items()}
    want = O.encoder_layer_forward(p64, "encoder.layers.0.", src.bfloat16().double(), O.encoder_reference_points(shapes, B).double(),
                                   shapes, torch.ones(B, Lv).double(), pos.bfloat16().double().expand(B, -1, -1))
    s16, p16 = src.to(cuda_dev).bfloat16(), pos.to(cuda_dev).bfloat16()
    fused = layer(s16, ref, shapes, None, p16)
    # unfused composition through the same public pieces
    pk = layer._packed_weights(torch.bfloat16)
    conv = ops.conv3x3_tokens(s16, pk["conv_w"], shapes)
    gn = ops.groupnorm_stats(conv, shapes, groups=32)
    src2 = layer.self_attn(ops.add_bcast(s16, p16), ref, s16, shapes, None)
    x = ops.residual_layernorm(src2, s16, pk["n1w"], pk["n1b"])
    h = ops.linear(x, pk["w1"], pk["b1"], w_transposed=True, epilogue=L.EPI_RELU)
    f = ops.linear(h, pk["w2"], pk["b2"], w_transposed=True)
    unfused = ops.residual_layernorm_gn(f, x, pk["n2w"], pk["n2b"], conv, s16, gn, pk["gn_w"], pk["gn_b"], shapes, groups=32)
    e_f, e_u = l2_err(fused.float(), want), l2_err(unfused.float(), want)
    assert e_f < 1e-2 and e_u < 1e-2, (e_f, e_u)
    assert l2_err(fused.float(), unfused.float().cpu()) < 1e-2


@pytest.mark.parametrize("B,C,H,W,N", [(3, 512, 16, 16, 256), (2, 1024, 16, 8, 256), (5, 2048, 8, 16, 256), (2, 256, 32, 32, 256),
                                       (1, 128, 16, 16, 128)])
def test_linear_reads_channel_major_feature_maps(cuda_dev, B, C, H, W, N):
    """x_nchw: the 1x1 convolution of input_proj (t_e_d.py:417-419) as a GEMM whose A operand is the NCHW map itself
    (MN-major tcgen05 operand) — bit-equal to transposing to tokens first: same products, same accumulation order."""
    rng = np.random.Generator(np.random.PCG64(B * C + N))
    f = torch.from_numpy(O.rng_normal(rng, (B, C, H, W))).to(cuda_dev).bfloat16()
    w = torch.from_numpy(O.rng_uniform(rng, (C, N), (6.0 / (C + N)) ** 0.5)).to(cuda_dev)
    b = torch.from_numpy(O.rng_uniform(rng, (N,), 0.1)).to(cuda_dev)
    wp = torch.empty((N, C), dtype=torch.bfloat16, device=cuda_dev)
    ops.pack_weight(w, wp)
    want = ops.linear(ops.nchw_to_tokens(f), wp, b, w_transposed=True)
    got = ops.linear(f, wp, b, w_transposed=True, x_nchw=True)
    assert got.shape == want.shape == (B, H * W, N)
    assert torch.equal(got, want)
    ref = f.double().flatten(2).transpose(1, 2) @ wp.double().T + b.double()
    assert rel_err(got.float(), ref.cpu()) < 8e-3
    with pytest.raises(L.EmrtError):
        ops.linear(f, wp, b, w_transposed=True, x_nchw=True, epilogue=L.EPI_RELU)
