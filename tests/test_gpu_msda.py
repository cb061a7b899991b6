"""GPU parity tests of the MSDA path: CUDA kernels (through the C ABI) vs the CPU oracle on the same seeded inputs.
Tolerances: fp32 1e-4 relative (of the output's max magnitude), bf16 1e-2 relative (BASELINE.json)."""
import os

import numpy as np
import pytest
import torch

import oracle as O
import emrt_b200
from emrt_b200 import ops, _lib as L

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
FP32_TOL, BF16_TOL = 1e-4, 1e-2


def rel_err(got, want):
    want = torch.as_tensor(want).double()
    return ((got.detach().double().cpu() - want).abs().max() / want.abs().max().clamp_min(1e-30)).item()


def _gather_case(seed, shapes, B, M, D, P, Lq, spread=0.3):
    rng = np.random.Generator(np.random.PCG64(seed))
    _, Lv = O.level_tables(shapes)
    L_ = len(shapes)
    value = O.rng_normal(rng, (B, Lv, M, D))
    loc = rng.uniform(-spread, 1 + spread, size=(B, Lq, M, L_, P, 2)).astype(np.float32)
    attn = rng.uniform(0, 1, size=(B, Lq, M, L_, P)).astype(np.float32)
    attn /= attn.reshape(B, Lq, M, -1).sum(-1)[..., None, None]
    return value, loc, attn


CASES = [
    # shapes, B, M, D, P, Lq
    ([(32, 32), (16, 16), (8, 8)], 2, 8, 32, 6, 1344),       # cfg 1 / cfg 2 geometry (encoder)
    ([(32, 32), (16, 16), (8, 8)], 3, 8, 32, 6, 110),        # decoder cross-attention, Lq != Lv
    ([(8, 6), (4, 3), (2, 2)], 1, 2, 32, 6, 37),             # non-square levels, B = 1
    ([(5, 9)], 2, 3, 16, 5, 13),                             # single level, odd P, M not a power of two, D = 16
    ([(7, 3), (2, 5)], 2, 4, 64, 3, 9),                      # D = 64
]


@pytest.mark.parametrize("shapes,B,M,D,P,Lq", CASES)
def test_gather_fwd_fp32_matches_oracle(cuda_dev, shapes, B, M, D, P, Lq):
    value, loc, attn = _gather_case(0, shapes, B, M, D, P, Lq)
    want = O.gather_corner_loop(value, shapes, loc, attn)
    got = emrt_b200.deformable_attention_core_func(torch.from_numpy(value).to(cuda_dev), shapes,
                                                   torch.from_numpy(loc).to(cuda_dev), torch.from_numpy(attn).to(cuda_dev))
    assert got.shape == (B, Lq, M * D)
    assert rel_err(got, want) < FP32_TOL
    lib_want = O.deformable_attention_core_func(value, shapes, loc, attn)    # the reference's own op composition
    assert rel_err(got, lib_want) < FP32_TOL


@pytest.mark.parametrize("shapes,B,M,D,P,Lq", CASES)
@pytest.mark.parametrize("loc_dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_gather_fwd_bf16_matches_oracle(cuda_dev, shapes, B, M, D, P, Lq, loc_dtype):
    value, loc, attn = _gather_case(1, shapes, B, M, D, P, Lq)
    v16 = torch.from_numpy(value).bfloat16()
    l16 = torch.from_numpy(loc).to(loc_dtype)
    a16 = torch.from_numpy(attn).to(loc_dtype)
    # oracle on exactly the rounded inputs, float64 arithmetic: only the output rounding differs
    want = O.gather_corner_loop(v16.float().numpy(), shapes, l16.float().numpy(), a16.float().numpy())
    got = emrt_b200.deformable_attention_core_func(v16.to(cuda_dev), shapes, l16.to(cuda_dev), a16.to(cuda_dev))
    assert got.dtype == torch.bfloat16
    assert rel_err(got.float(), want) < BF16_TOL


def test_gather_pixel_offset_mode_equals_normalized_mode(cuda_dev):
    shapes = [(32, 32), (16, 16), (8, 8)]
    B, M, D, P = 2, 8, 32, 6
    rng = np.random.Generator(np.random.PCG64(3))
    _, Lv = O.level_tables(shapes)
    Lq = Lv
    value = O.rng_normal(rng, (B, Lv, M, D))
    off = O.rng_normal(rng, (B, Lq, M, 3, P, 2), 3.0)
    attn = rng.uniform(0, 1, size=(B, Lq, M, 3, P)).astype(np.float32)
    ref = O.encoder_reference_points(shapes, B).numpy()
    norm = np.array([[w, h] for h, w in shapes], np.float32).reshape(1, 1, 1, 3, 1, 2)
    loc = ref.reshape(B, Lq, 1, 3, 1, 2) + off / norm
    want = O.gather_corner_loop(value, shapes, loc, attn)
    d = lambda a: torch.from_numpy(a).to(cuda_dev)
    got = ops.msda_gather_fwd(d(value), d(off), d(attn), shapes, ref=d(ref), mode=L.LOC_PIXEL_OFFSET)
    assert rel_err(got, want) < FP32_TOL
    got_shared = ops.msda_gather_fwd(d(value), d(off), d(attn), shapes, ref=d(ref[:1]), mode=L.LOC_PIXEL_OFFSET)
    assert torch.equal(got, got_shared)          # batch-shared reference points (stride 0)
    # bf16 value + fp16 pixel offsets (the fused path's storage format)
    v16, o16, a16 = d(value).bfloat16(), d(off).half(), d(attn).half()
    loc16 = ref.reshape(B, Lq, 1, 3, 1, 2) + o16.float().cpu().numpy() / norm
    want16 = O.gather_corner_loop(v16.float().cpu().numpy(), shapes, loc16, a16.float().cpu().numpy())
    got16 = ops.msda_gather_fwd(v16, o16, a16, shapes, ref=d(ref), mode=L.LOC_PIXEL_OFFSET)
    assert rel_err(got16.float(), want16) < BF16_TOL


def test_gather_properties(cuda_dev):
    shapes = [(16, 12), (8, 6)]
    value, loc, attn = _gather_case(5, shapes, 2, 4, 32, 4, 50)
    d = lambda a: torch.from_numpy(a).to(cuda_dev)
    f = lambda v, l=loc, a=attn: emrt_b200.deformable_attention_core_func(d(v), shapes, d(l), d(a))
    # all samples outside -> exact zero
    assert torch.all(f(value, loc * 0 + 2.5) == 0)
    assert torch.all(f(value, loc * 0 - 1.5) == 0)
    # linearity in value
    v2 = np.roll(value, 3, axis=1)
    assert rel_err(f(2 * value + 3 * v2), (2 * f(value) + 3 * f(v2)).cpu()) < 1e-5
    # permuting heads permutes output channel blocks
    perm = [2, 0, 3, 1]
    out = f(value).reshape(2, 50, 4, 32)
    outp = f(value[:, :, perm], loc[:, :, perm], attn[:, :, perm]).reshape(2, 50, 4, 32)
    assert torch.equal(outp, out[:, :, perm])
    # NaN / huge coordinates contribute nothing (guarded float->int conversion)
    loc_bad = loc.copy()
    loc_bad[0, 0, 0, 0, 0] = [1e30, -1e30]
    assert torch.isfinite(f(value, loc_bad)).all()


@pytest.mark.parametrize("mode", ["normalized", "pixel"])
def test_gather_bwd_matches_autograd_of_oracle(cuda_dev, mode):
    shapes = [(12, 10), (6, 5), (3, 3)]
    B, M, D, P, Lq = 2, 4, 32, 6, 40
    value, loc, attn = _gather_case(7, shapes, B, M, D, P, Lq, spread=0.2)
    rng = np.random.Generator(np.random.PCG64(8))
    gout = O.rng_normal(rng, (B, Lq, M * D))
    tv = torch.from_numpy(value).double().requires_grad_()
    tl = torch.from_numpy(loc).double().requires_grad_()
    ta = torch.from_numpy(attn).double().requires_grad_()
    O.deformable_attention_core_func(tv, shapes, tl, ta).backward(torch.from_numpy(gout).double())
    d = lambda a: torch.from_numpy(a).to(cuda_dev)
    if mode == "normalized":
        gv, gl, ga = ops.msda_gather_bwd(d(gout), d(value), d(loc), d(attn), shapes)
        want_gl = tl.grad
    else:
        # same sample positions expressed as pixel offsets from random reference points
        ref = rng.uniform(0, 1, size=(B, Lq, len(shapes), 2)).astype(np.float32)
        norm = np.array([[w, h] for h, w in shapes], np.float64).reshape(1, 1, 1, -1, 1, 2)
        off = ((loc.astype(np.float64) - ref.reshape(B, Lq, 1, -1, 1, 2)) * norm).astype(np.float32)
        gv, gl, ga = ops.msda_gather_bwd(d(gout), d(value), d(off), d(attn), shapes, ref=d(ref), mode=L.LOC_PIXEL_OFFSET)
        want_gl = tl.grad / torch.from_numpy(norm)       # d/d off_px = d/d loc / (W, H)
    tol = 2e-4 if mode == "normalized" else 2e-3         # pixel mode re-derives positions in fp32
    assert rel_err(gv, tv.grad) < tol
    assert rel_err(ga, ta.grad) < tol
    assert rel_err(gl, want_gl) < tol


@pytest.mark.parametrize("rows,K,N,wt", [(300, 256, 256, False), (129, 256, 432, True), (64, 96, 40, False),
                                         (1000, 1024, 256, True)])
def test_linear_simt_fp32(cuda_dev, rows, K, N, wt):
    rng = np.random.Generator(np.random.PCG64(9))
    x, w, b = O.rng_normal(rng, (rows, K)), O.rng_normal(rng, (K, N), 0.1), O.rng_normal(rng, (N,))
    scale = rng.uniform(0, 1, size=(rows,)).astype(np.float32)
    want = np.maximum((x.astype(np.float64) @ w.astype(np.float64) + b) * scale[:, None], 0)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda_dev)
    got = ops.linear(d(x), d(w.T if wt else w), d(b), w_transposed=wt, epilogue=L.EPI_ROW_MASK | L.EPI_RELU,
                     row_scale=d(scale), impl=L.IMPL_SIMT)
    assert rel_err(got, want) < 1e-5


def _load_module(params, C, M, Lv, P, dev, train=False):
    m = emrt_b200.MSDeformableAttention(C, M, Lv, P).to(dev)
    with torch.no_grad():
        for name, arr in params.items():
            mod, leaf = name.split(".")
            getattr(getattr(m, mod), leaf).copy_(torch.from_numpy(arr))
    m.requires_grad_(train)        # frozen parameters -> the inference path; trainable -> the autograd path
    return m


def test_msda_module_fp32_golden(cuda_dev):
    g = np.load(os.path.join(GOLD, "msda_small.npz"))
    shapes = [tuple(s) for s in g["shapes"].tolist()]
    params = {k[2:]: g[k] for k in g.files if k.startswith("p.")}
    m = _load_module(params, 64, 2, 3, 6, cuda_dev)
    d = lambda a: torch.from_numpy(a).to(cuda_dev)
    got = m(d(g["query"]), d(g["ref"]), d(g["value"]), torch.tensor(shapes), d(g["mask"]))
    assert rel_err(got, g["out"]) < FP32_TOL


@pytest.mark.parametrize("Lq_mode", ["encoder", "decoder"])
def test_msda_module_fp32_cfg1_shape(cuda_dev, Lq_mode):
    shapes = [(32, 32), (16, 16), (8, 8)]
    B, C, M, P = 2, 256, 8, 6
    rng = np.random.Generator(np.random.PCG64(0))
    _, Lv = O.level_tables(shapes)
    params = O.make_msda_params(1234, C, M, 3, P)
    v = O.rng_normal(rng, (B, Lv, C))
    if Lq_mode == "encoder":
        q, ref = O.rng_normal(rng, (B, Lv, C)), O.encoder_reference_points(shapes, B).numpy()
    else:
        q = O.rng_normal(rng, (B, 110, C))
        ref = np.repeat(rng.uniform(0, 1, size=(B, 110, 1, 2)).astype(np.float32), 3, axis=2)
    want = O.msda_forward(params, q, ref, v, shapes, None, M, P, dtype=torch.float64)
    m = _load_module(params, C, M, 3, P, cuda_dev)
    d = lambda a: torch.from_numpy(a).to(cuda_dev)
    got = m(d(q), d(ref), d(v), shapes)
    assert rel_err(got, want) < FP32_TOL


@pytest.mark.parametrize("impl", ["simt", "tcgen05"])
def test_msda_module_bf16(cuda_dev, impl):
    shapes = [(32, 32), (16, 16), (8, 8)]
    B, C, M, P = 4, 256, 8, 6
    rng = np.random.Generator(np.random.PCG64(1))
    _, Lv = O.level_tables(shapes)
    params = O.make_msda_params(1234, C, M, 3, P)
    q, v = O.rng_normal(rng, (B, Lv, C)), O.rng_normal(rng, (B, Lv, C))
    ref = O.encoder_reference_points(shapes, B).numpy()
    mask = (rng.uniform(size=(B, Lv)) > 0.05).astype(np.float32)
    # oracle in float64 on the bf16-rounded inputs and weights
    r = lambda a: torch.from_numpy(a).bfloat16().float().numpy()
    p16 = {k: (r(a) if k.endswith("weight") else a) for k, a in params.items()}
    want = O.msda_forward(p16, r(q), ref, r(v), shapes, mask, M, P, dtype=torch.float64)
    m = _load_module(params, C, M, 3, P, cuda_dev)
    m.gemm_impl = L.IMPL_SIMT if impl == "simt" else L.IMPL_TCGEN05
    d = lambda a: torch.from_numpy(a).to(cuda_dev)
    got = m(d(q).bfloat16(), d(ref), d(v).bfloat16(), shapes, d(mask))
    assert got.dtype == torch.bfloat16
    assert rel_err(got.float(), want) < BF16_TOL


@pytest.mark.parametrize("rows,K,N,epi", [
    (300, 256, 256, "none"), (128, 256, 256, "mask"), (1000, 256, 432, "qproj"), (129, 1024, 256, "none"),
    (77, 256, 1024, "relu"), (5000, 256, 64, "none"), (260, 64, 128, "mask"), (20000, 256, 256, "none"),
    (96768, 256, 432, "qproj"),
])
def test_linear_tcgen05_matches_fp64(cuda_dev, rows, K, N, epi):
    """tcgen05/TMEM/TMA GEMM vs float64 on the same bf16-rounded operands (fp32 accumulation: ~1e-6 relative)."""
    rng = np.random.Generator(np.random.PCG64(rows + N))
    x = torch.from_numpy(O.rng_normal(rng, (rows, K))).bfloat16()
    w = torch.from_numpy(O.rng_normal(rng, (K, N), 0.1)).bfloat16()          # Paddle layout [in, out]
    b = torch.from_numpy(O.rng_normal(rng, (N,)))
    wt = torch.empty((N, K), dtype=torch.bfloat16, device=cuda_dev)
    ops.pack_weight(w.to(cuda_dev), wt)
    assert torch.equal(wt.cpu(), w.t().contiguous())
    ref = x.double() @ w.double() + b.double()
    xd, bd = x.to(cuda_dev), b.to(cuda_dev)
    if epi == "qproj":
        tp = N // 3
        off, attn = ops.linear(xd, wt, bd, w_transposed=True, y_dtype=torch.float16, epilogue=L.EPI_MSDA_QPROJ,
                               qproj_group=18, impl=L.IMPL_TCGEN05)
        assert off.shape == (rows, 2 * tp) and attn.shape == (rows, tp)
        assert rel_err(off.float(), ref[:, :2 * tp]) < 2e-3                   # fp16 storage of the offsets
        want = torch.softmax(ref[:, 2 * tp:].reshape(rows, -1, 18), -1).reshape(rows, tp)
        assert (attn.float().cpu().double() - want).abs().max() < 2e-3
        assert (attn.float().reshape(rows, -1, 18).sum(-1) - 1).abs().max() < 5e-3
        return
    scale = torch.from_numpy(rng.uniform(0, 1, size=(rows,)).astype(np.float32))
    flags, kw = L.EPI_NONE, {}
    if epi == "mask":
        flags, kw, ref = L.EPI_ROW_MASK, dict(row_scale=scale.to(cuda_dev)), ref * scale.double()[:, None]
    if epi == "relu":
        flags, ref = L.EPI_RELU, ref.clamp_min(0)
    got32 = ops.linear(xd, wt, bd, w_transposed=True, y_dtype=torch.float32, epilogue=flags, impl=L.IMPL_TCGEN05, **kw)
    assert rel_err(got32, ref) < 1e-5
    got16 = ops.linear(xd, wt, bd, w_transposed=True, epilogue=flags, impl=L.IMPL_TCGEN05, **kw)
    assert got16.dtype == torch.bfloat16 and rel_err(got16.float(), ref) < 1e-2
    simt = ops.linear(xd, wt, bd, w_transposed=True, y_dtype=torch.float32, epilogue=flags, impl=L.IMPL_SIMT, **kw)
    assert rel_err(got32, simt.cpu()) < 1e-5


@pytest.mark.parametrize("Lq,B", [(1344, 2), (110, 3), (5376, 1), (77, 1)])
@pytest.mark.parametrize("loc_dtype", [torch.float16, torch.float32])
def test_gather_v1_specialised_kernel_all_layouts(cuda_dev, Lq, B, loc_dtype):
    """bf16 / D=32 / 3x6 specialised kernel: pixel-major and head-major value layouts, both loc modes, against the
    float64 oracle and against the generic kernel (EMRT_GATHER_V0=1)."""
    shapes = [(32, 32), (16, 16), (8, 8)] if Lq != 5376 else [(64, 64), (32, 32), (16, 16)]
    M, D, P = 8, 32, 6
    rng = np.random.Generator(np.random.PCG64(Lq))
    _, Lv = O.level_tables(shapes)
    value = torch.from_numpy(O.rng_normal(rng, (B, Lv, M, D))).bfloat16()
    off = torch.from_numpy(O.rng_normal(rng, (B, Lq, M, 3, P, 2), 4.0)).to(loc_dtype)
    attn = torch.from_numpy(rng.uniform(0, 1, size=(B, Lq, M, 3, P)).astype(np.float32)).to(loc_dtype)
    ref = rng.uniform(0, 1, size=(B, Lq, 3, 2)).astype(np.float32)
    norm = np.array([[w, h] for h, w in shapes], np.float32).reshape(1, 1, 1, 3, 1, 2)
    loc = ref.reshape(B, Lq, 1, 3, 1, 2) + off.float().numpy() / norm
    want = O.gather_corner_loop(value.float().numpy(), shapes, loc, attn.float().numpy())
    d = lambda t: (torch.from_numpy(t) if isinstance(t, np.ndarray) else t).to(cuda_dev)
    vd, od, ad, rd = d(value), d(off), d(attn), d(ref)
    v_hm = vd.permute(0, 2, 1, 3).contiguous()
    got_pm = ops.msda_gather_fwd(vd, od, ad, shapes, ref=rd, mode=L.LOC_PIXEL_OFFSET)
    got_hm = ops.msda_gather_fwd(v_hm, od, ad, shapes, ref=rd, mode=L.LOC_PIXEL_OFFSET | L.VALUE_HEAD_MAJOR)
    os.environ["EMRT_GATHER_V0"] = "1"
    try:
        got_v0 = ops.msda_gather_fwd(vd, od, ad, shapes, ref=rd, mode=L.LOC_PIXEL_OFFSET)
    finally:
        del os.environ["EMRT_GATHER_V0"]
    assert rel_err(got_pm.float(), want) < BF16_TOL
    assert torch.equal(got_pm, got_hm)                      # same arithmetic, different value layout
    assert rel_err(got_pm.float(), got_v0.float().cpu()) < BF16_TOL
    # normalised-location mode through the same kernel
    locd = d(loc.astype(np.float32)).to(loc_dtype)
    want_n = O.gather_corner_loop(value.float().numpy(), shapes, locd.float().cpu().numpy(), attn.float().numpy())
    got_n = ops.msda_gather_fwd(v_hm, locd, ad, shapes, mode=L.LOC_NORMALIZED | L.VALUE_HEAD_MAJOR)
    assert rel_err(got_n.float(), want_n) < BF16_TOL


@pytest.mark.parametrize("tile,B,spread,R", [(512, 2, 3.0, None), (256, 3, 3.0, None), (512, 1, 9.0, None),
                                              (256, 2, 3.0, "2"), (128, 2, 2.0, None)])
@pytest.mark.parametrize("loc_dtype", [torch.float16, torch.float32])
def test_gather_window_staged_kernel(cuda_dev, tile, B, spread, R, loc_dtype):
    """EMRT_QUERY_PIXEL_GRID (TMA-staged value windows, fma.rn.f32.bf16): encoder geometry, pixel-centre reference
    points, offsets = reference-init directions + noise.  Checked against the float64 oracle and the L1-path kernel.
    spread 9 / R=2 push many samples out of the staged window (slow path) and out of the map (zero padding)."""
    shapes = [(tile // 8,) * 2, (tile // 16,) * 2, (tile // 32,) * 2]
    M, D, P = 8, 32, 6
    rng = np.random.Generator(np.random.PCG64(tile + B))
    _, Lv = O.level_tables(shapes)
    Lq = Lv
    value = torch.from_numpy(O.rng_normal(rng, (B, Lv, M, D))).bfloat16()
    bias = O.msda_reset_parameters(M * D, M, 3, P).reshape(1, 1, M, 3, P, 2)
    off = torch.from_numpy(bias + O.rng_normal(rng, (B, Lq, M, 3, P, 2), spread)).to(loc_dtype)
    attn = torch.from_numpy(rng.uniform(0, 1, size=(B, Lq, M, 3, P)).astype(np.float32)).to(loc_dtype)
    ref_t = emrt_b200.get_reference_points(shapes, device=cuda_dev)
    assert getattr(ref_t, "pixel_grid", False) and tuple(ref_t.shape) == (1, Lv, 3, 2)
    assert torch.equal(ref_t.cpu(), O.encoder_reference_points(shapes, 1))
    ref = ref_t.cpu().numpy()
    norm = np.array([[w, h] for h, w in shapes], np.float32).reshape(1, 1, 1, 3, 1, 2)
    loc = ref.reshape(1, Lq, 1, 3, 1, 2) + off.float().numpy() / norm
    want = O.gather_corner_loop(value.float().numpy(), shapes, loc, attn.float().numpy())
    d = lambda t: (torch.from_numpy(t) if isinstance(t, np.ndarray) else t).to(cuda_dev)
    v_hm, od, ad = d(value).permute(0, 2, 1, 3).contiguous(), d(off), d(attn)
    base = L.LOC_PIXEL_OFFSET | L.VALUE_HEAD_MAJOR
    got_l1 = ops.msda_gather_fwd(v_hm, od, ad, shapes, ref=ref_t, mode=base)
    if R is not None:
        os.environ["EMRT_WIN_R"] = R
    try:
        before = ops.launch_count()
        got = ops.msda_gather_fwd(v_hm, od, ad, shapes, ref=ref_t, mode=base | L.QUERY_PIXEL_GRID)
        assert ops.launch_count() == before + 1
        # normalised-location mode through the same kernel
        locd = d(loc.astype(np.float32)).to(loc_dtype)
        got_n = ops.msda_gather_fwd(v_hm, locd, ad, shapes, mode=L.LOC_NORMALIZED | L.VALUE_HEAD_MAJOR | L.QUERY_PIXEL_GRID)
        # window-centre hints (emrt_msda_gather_fwd_hint) only move the staged windows: the matching hint (mid-range of
        # the offset bias per head and level) and a wild one (every sample leaves its window) give the same result
        b4 = bias.reshape(M, 3, P, 2)
        mid = np.rint((b4.max(axis=2) + b4.min(axis=2)) * 0.5).astype(np.int32)
        got_hint = ops.msda_gather_fwd(v_hm, od, ad, shapes, ref=ref_t, mode=base | L.QUERY_PIXEL_GRID,
                                       win_center=L.i32_array(mid.reshape(-1).tolist()))
        wild = np.where(np.arange(M * 3 * 2) % 2 == 0, 40, -40).astype(np.int32)
        got_wild = ops.msda_gather_fwd(v_hm, od, ad, shapes, ref=ref_t, mode=base | L.QUERY_PIXEL_GRID,
                                       win_center=L.i32_array(wild.tolist()))
        # the reference's own pixel-major value layout through the same kernel (5-D tensor maps): same arithmetic
        got_pm = ops.msda_gather_fwd(d(value), od, ad, shapes, ref=ref_t, mode=L.LOC_PIXEL_OFFSET | L.QUERY_PIXEL_GRID)
        assert ops.launch_count() == before + 5
    finally:
        os.environ.pop("EMRT_WIN_R", None)
    torch.cuda.synchronize()
    assert torch.equal(got_pm, got)
    assert rel_err(got.float(), want) < BF16_TOL
    assert rel_err(got_hint.float(), want) < BF16_TOL
    assert rel_err(got_wild.float(), want) < BF16_TOL
    assert rel_err(got.float(), got_l1.float().cpu()) < BF16_TOL
    want_n = O.gather_corner_loop(value.float().numpy(), shapes, locd.float().cpu().numpy(), attn.float().numpy())
    assert rel_err(got_n.float(), want_n) < BF16_TOL


@pytest.mark.parametrize("version", ["v2", "v1"])
@pytest.mark.parametrize("tile,B,spread,R,loc_dtype", [(256, 2, 1.0, None, torch.float16), (512, 1, 1.0, None, torch.float16),
                                                        (128, 2, 3.0, None, torch.float32), (256, 1, 3.0, "2", torch.float16)])
def test_gather_bwd_windowed_kernel(cuda_dev, tile, B, spread, R, loc_dtype, version):
    """EMRT_QUERY_PIXEL_GRID backward: grad_value accumulated in fixed point in shared-memory windows (integer
    shared-memory reductions), grad_loc / grad_attn by shuffles.  Checked against the generic backward (float
    reductions in L2) on the same bf16 inputs, against torch autograd through the float64 oracle, with and without a
    window-centre hint, and in both location modes.  spread 3 / R=2 push samples out of the windows (global fallback)
    and out of the map."""
    shapes = [(tile // 8,) * 2, (tile // 16,) * 2, (tile // 32,) * 2]
    M, D, P = 8, 32, 6
    rng = np.random.Generator(np.random.PCG64(100 + tile + B))
    _, Lv = O.level_tables(shapes)
    Lq = Lv
    value = torch.from_numpy(O.rng_normal(rng, (B, Lv, M, D))).bfloat16()
    gout = torch.from_numpy(O.rng_normal(rng, (B, Lq, M * D), 3.0)).bfloat16()
    bias = O.msda_reset_parameters(M * D, M, 3, P).reshape(1, 1, M, 3, P, 2)
    off = torch.from_numpy(bias + O.rng_normal(rng, (B, Lq, M, 3, P, 2), spread)).to(loc_dtype)
    attn = torch.from_numpy(rng.uniform(0, 1, size=(B, Lq, M, 3, P)).astype(np.float32)).to(loc_dtype)
    ref_t = emrt_b200.get_reference_points(shapes, device=cuda_dev)
    d = lambda t: t.to(cuda_dev)
    vd, gd, od, ad = d(value), d(gout), d(off), d(attn)
    base = L.LOC_PIXEL_OFFSET
    want = ops.msda_gather_bwd(gd, vd, od, ad, shapes, ref=ref_t, mode=base)
    b4 = bias.reshape(M, 3, P, 2)
    mid = np.rint((b4.max(axis=2) + b4.min(axis=2)) * 0.5).astype(np.int32)
    # v2 (default): TMA-staged value windows + footprint records, one CTA per SM; v1: corners through L1, two CTAs per SM
    # (what runs when v2's windows do not fit one CTA's shared memory)
    if R is not None:
        os.environ["EMRT_BWD_WIN_R"] = R
    if version == "v1":
        os.environ["EMRT_BWD_WIN_V1"] = "1"
    try:
        before = ops.launch_count()
        got = ops.msda_gather_bwd(gd, vd, od, ad, shapes, ref=ref_t, mode=base | L.QUERY_PIXEL_GRID)
        got_hint = ops.msda_gather_bwd(gd, vd, od, ad, shapes, ref=ref_t, mode=base | L.QUERY_PIXEL_GRID,
                                       win_center=L.i32_array(mid.reshape(-1).tolist()))
        # normalised locations through the same kernel
        norm = np.array([[w, h] for h, w in shapes], np.float32).reshape(1, 1, 1, 3, 1, 2)
        loc = (ref_t.cpu().numpy().reshape(1, Lq, 1, 3, 1, 2) + off.float().numpy() / norm).astype(np.float32)
        locd = d(torch.from_numpy(loc))
        ad32 = ad.float()
        want_n = ops.msda_gather_bwd(gd, vd, locd, ad32, shapes, mode=L.LOC_NORMALIZED)
        got_n = ops.msda_gather_bwd(gd, vd, locd, ad32, shapes, mode=L.LOC_NORMALIZED | L.QUERY_PIXEL_GRID)
    finally:
        os.environ.pop("EMRT_BWD_WIN_R", None)
        os.environ.pop("EMRT_BWD_WIN_V1", None)
    torch.cuda.synchronize()
    for name, g, w in [("plain", got, want), ("hint", got_hint, want), ("normalized", got_n, want_n)]:
        # grad_value: fixed point with 2^-21 of the CTA's max |grad_out| per contribution; grad_loc / grad_attn: the same
        # bf16 products summed in another order
        assert rel_err(g[0], w[0].cpu()) < 2e-5, name
        assert rel_err(g[1], w[1].cpu()) < 2e-5, name
        assert rel_err(g[2], w[2].cpu()) < 2e-5, name
    # and against autograd through the oracle on the same (bf16 / fp16-rounded) inputs
    tv = value.double().requires_grad_()
    tl = torch.from_numpy(loc).double().requires_grad_()
    ta = attn.double().requires_grad_()
    O.deformable_attention_core_func(tv, shapes, tl, ta).backward(gout.double().view(B, Lq, M * D))
    # (a sample within ~1e-6 px of a pixel boundary lands in the neighbouring bilinear cell of the float64 oracle, where
    # its own gradients differ by O(1): judge grad_loc / grad_attn by the fraction of such elements)
    assert rel_err(got_n[0], tv.grad) < 2e-4
    for g, w in [(got_n[2], ta.grad), (got_n[1], tl.grad)]:
        bad = ((g.double().cpu() - w).abs() > 2e-3 * w.abs().max()).double().mean().item()
        assert bad < 1e-4, bad


def test_gather_window_staged_arbitrary_reference_points(cuda_dev):
    """The pixel-grid flag is only a locality promise: random reference points (almost every sample leaves its
    region's window) must still give the oracle's result."""
    shapes = [(32, 32), (16, 16), (8, 8)]
    M, D, P, B = 8, 32, 6, 1
    rng = np.random.Generator(np.random.PCG64(5))
    _, Lv = O.level_tables(shapes)
    value = torch.from_numpy(O.rng_normal(rng, (B, Lv, M, D))).bfloat16()
    loc = rng.uniform(-0.2, 1.2, size=(B, Lv, M, 3, P, 2)).astype(np.float32)
    attn = rng.uniform(0, 1, size=(B, Lv, M, 3, P)).astype(np.float32)
    want = O.gather_corner_loop(value.float().numpy(), shapes, loc, attn)
    v_hm = value.to(cuda_dev).permute(0, 2, 1, 3).contiguous()
    got = ops.msda_gather_fwd(v_hm, torch.from_numpy(loc).to(cuda_dev), torch.from_numpy(attn).to(cuda_dev), shapes,
                              mode=L.LOC_NORMALIZED | L.VALUE_HEAD_MAJOR | L.QUERY_PIXEL_GRID)
    assert rel_err(got.float(), want) < BF16_TOL


def _oracle_msda_grads(params, q, ref, v, shapes, mask, M, P, d_out, dtype=torch.float64):
    """Gradients of the oracle's MSDA forward w.r.t. query, value and the eight parameters (torch autograd through
    the grid_sample formulation — what Paddle autograd derives for the reference module)."""
    tp = {k: torch.from_numpy(a).to(dtype).requires_grad_(True) for k, a in params.items()}
    tq = torch.from_numpy(q).to(dtype).requires_grad_(True)
    tv = torch.from_numpy(v).to(dtype).requires_grad_(True)
    out = O.msda_forward(tp, tq, torch.from_numpy(ref).to(dtype), tv, shapes,
                         None if mask is None else torch.from_numpy(mask).to(dtype), M, P, dtype=dtype)
    out.backward(torch.from_numpy(d_out).to(dtype))
    grads = {k: t.grad for k, t in tp.items()}
    grads["query"], grads["value"] = tq.grad, tv.grad
    return out.detach(), grads


@pytest.mark.parametrize("case", ["encoder", "decoder"])
def test_msda_module_backward_fp32(cuda_dev, case):
    """Training path, fp32: fwd + bwd through the autograd Function vs autograd of the oracle (1e-4)."""
    shapes = [(16, 16), (8, 8), (4, 4)]
    B, C, M, P = 2, 256, 8, 6
    rng = np.random.Generator(np.random.PCG64(11))
    _, Lv = O.level_tables(shapes)
    Lq = Lv if case == "encoder" else 37
    params = O.make_msda_params(77, C, M, 3, P)
    q, v = O.rng_normal(rng, (B, Lq, C)), O.rng_normal(rng, (B, Lv, C))
    ref = (O.encoder_reference_points(shapes, B).numpy() if case == "encoder"
           else rng.uniform(0.1, 0.9, size=(B, Lq, 3, 2)).astype(np.float32))
    mask = (rng.uniform(size=(B, Lv)) > 0.1).astype(np.float32)
    d_out = O.rng_normal(rng, (B, Lq, C))
    want, wg = _oracle_msda_grads(params, q, ref, v, shapes, mask, M, P, d_out)
    m = _load_module(params, C, M, 3, P, cuda_dev, train=True)
    d = lambda a: torch.from_numpy(a).to(cuda_dev)
    tq, tv = d(q).requires_grad_(True), d(v).requires_grad_(True)
    out = m(tq, d(ref), tv, shapes, d(mask))
    out.backward(d(d_out))
    assert rel_err(out, want) < FP32_TOL
    assert rel_err(tq.grad, wg["query"]) < 5e-4 and rel_err(tv.grad, wg["value"]) < 5e-4
    for name, g in wg.items():
        if "." in name:
            mod, leaf = name.split(".")
            got = getattr(getattr(m, mod), leaf).grad
            assert got is not None and rel_err(got, g) < 5e-4, name


@pytest.mark.parametrize("impl", ["simt", "tcgen05", "tcgen05-windowed"])
def test_msda_module_backward_bf16(cuda_dev, impl):
    """Training path, bf16 activations: gradients vs the float64 oracle on the bf16-rounded inputs / weights.
    "tcgen05-windowed": reference points from get_reference_points (tagged pixel grid), so the backward gather is the
    windowed kernel (integer shared-memory accumulation) with the module's window-centre hint."""
    shapes = [(32, 32), (16, 16), (8, 8)]
    B, C, M, P = 2, 256, 8, 6
    rng = np.random.Generator(np.random.PCG64(12))
    _, Lv = O.level_tables(shapes)
    params = O.make_msda_params(78, C, M, 3, P)
    q, v = O.rng_normal(rng, (B, Lv, C)), O.rng_normal(rng, (B, Lv, C))
    ref = O.encoder_reference_points(shapes, B).numpy()
    mask = (rng.uniform(size=(B, Lv)) > 0.1).astype(np.float32)
    d_out = O.rng_normal(rng, (B, Lv, C))
    r = lambda a: torch.from_numpy(a).bfloat16().float().numpy()
    p16 = {k: (r(a) if k.endswith("weight") else a) for k, a in params.items()}
    want, wg = _oracle_msda_grads(p16, r(q), ref, r(v), shapes, mask, M, P, r(d_out))
    m = _load_module(params, C, M, 3, P, cuda_dev, train=True)
    m.gemm_impl = L.IMPL_SIMT if impl == "simt" else L.IMPL_TCGEN05
    d = lambda a: torch.from_numpy(a).to(cuda_dev)
    tq, tv = d(q).bfloat16().requires_grad_(True), d(v).bfloat16().requires_grad_(True)
    if impl.endswith("windowed"):
        ref_t = emrt_b200.get_reference_points(shapes, device=cuda_dev)
        assert getattr(ref_t, "pixel_grid", False) and torch.equal(ref_t.cpu()[0], torch.from_numpy(ref)[0])
        m.packed_weights()                              # keep the one-off weight packing out of the launch counts
        before = ops.launch_count()
        out = m(tq, ref_t, tv, shapes, d(mask))
        out.backward(d(d_out).bfloat16())
        n_win = ops.launch_count() - before
        os.environ["EMRT_GATHER_NO_WIN"] = "1"          # same call through the generic backward: same launch count
        try:
            tq2, tv2 = d(q).bfloat16().requires_grad_(True), d(v).bfloat16().requires_grad_(True)
            m2 = _load_module(params, C, M, 3, P, cuda_dev, train=True)     # parameter .grad accumulates: a fresh copy
            m2.gemm_impl = m.gemm_impl
            m2.packed_weights()
            before = ops.launch_count()
            m2(tq2, ref_t, tv2, shapes, d(mask)).backward(d(d_out).bfloat16())
            assert ops.launch_count() - before == n_win
        finally:
            del os.environ["EMRT_GATHER_NO_WIN"]
        assert rel_err(tv.grad.float(), tv2.grad.float().cpu()) < 1e-2      # bf16 gradient tensors, different summation
    else:
        out = m(tq, d(ref), tv, shapes, d(mask))
        out.backward(d(d_out).bfloat16())
    # bf16 activations, fp16 offsets / weights and bf16 gradient tensors between the kernels.  The gradient w.r.t. a
    # sampling position is DISCONTINUOUS at pixel boundaries, and fp16-rounded offsets put ~1 % of the samples in the
    # neighbouring bilinear cell of the float64 oracle, so everything downstream of grad_loc (query, sampling_offsets)
    # is judged by its relative L2 error; the rest by the max-norm error.
    def l2_err(got, want):
        want = torch.as_tensor(want).double()
        return ((got.detach().double().cpu() - want).norm() / want.norm().clamp_min(1e-30)).item()
    errs = {"out": rel_err(out.float(), want), "value": rel_err(tv.grad.float(), wg["value"]),
            "query(l2)": l2_err(tq.grad.float(), wg["query"])}
    for name, g in wg.items():
        if "." in name:
            mod, leaf = name.split(".")
            got = getattr(getattr(m, mod), leaf).grad
            assert got is not None, name
            errs[name + ("(l2)" if mod == "sampling_offsets" else "")] = (
                l2_err(got.float(), g) if mod == "sampling_offsets" else rel_err(got.float(), g))
    assert errs["out"] < BF16_TOL, errs
    bad = {k: v for k, v in errs.items() if v >= (1e-1 if k.endswith("(l2)") else 3e-2)}
    assert not bad, errs


@pytest.mark.parametrize("mode", ["pixel", "normalized"])
@pytest.mark.parametrize("adt", [torch.float16, torch.bfloat16])
def test_msda_qproj_bwd_vectorised(cuda_dev, mode, adt):
    """Softmax + location backward (t_e_d.py:95,98-102): the vectorised kernel (16-bit weights, bf16 dq) against the
    element-wise one and against the closed form in float64."""
    shapes = [(16, 16), (8, 8), (4, 4)]
    M, P, B = 8, 6, 3
    _, Lv = O.level_tables(shapes)
    rng = np.random.Generator(np.random.PCG64(21))
    gl = torch.from_numpy(O.rng_normal(rng, (B, Lv, M, 3, P, 2))).to(cuda_dev)
    ga = torch.from_numpy(O.rng_normal(rng, (B, Lv, M, 3, P))).to(cuda_dev)
    attn = torch.softmax(torch.from_numpy(O.rng_normal(rng, (B, Lv, M, 3 * P))), -1).view(B, Lv, M, 3, P).to(adt).to(cuda_dev)
    md = L.LOC_PIXEL_OFFSET if mode == "pixel" else L.LOC_NORMALIZED
    before = ops.launch_count()
    got = ops.msda_qproj_bwd(gl, ga, attn, shapes, M, P, out_dtype=torch.bfloat16, mode=md)
    os.environ["EMRT_QPROJ_BWD_SLOW"] = "1"
    try:
        ref_k = ops.msda_qproj_bwd(gl, ga, attn, shapes, M, P, out_dtype=torch.bfloat16, mode=md)
    finally:
        del os.environ["EMRT_QPROJ_BWD_SLOW"]
    assert ops.launch_count() == before + 2
    assert torch.equal(got, ref_k)                         # same arithmetic, only the access pattern differs
    a64, g64 = attn.double().cpu().view(B, Lv, M, 18), ga.double().cpu().view(B, Lv, M, 18)
    dlogit = a64 * (g64 - (a64 * g64).sum(-1, keepdim=True))
    scale = torch.ones(3, 1, 2, dtype=torch.float64)
    if mode == "normalized":
        scale = torch.tensor([[1.0 / w, 1.0 / h] for h, w in shapes], dtype=torch.float64).view(3, 1, 2)
    doff = gl.double().cpu() * scale
    want = torch.cat([doff.reshape(B, Lv, -1), dlogit.reshape(B, Lv, -1)], -1)
    assert rel_err(got.float(), want) < BF16_TOL


@pytest.mark.parametrize("rows,K,N", [(1000, 256, 432), (129, 96, 40), (20000, 256, 256), (64, 256, 1024)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_linear_bwd_weight(cuda_dev, rows, K, N, dtype):
    """dW += x^T dy, db += colsum(dy): fp32 SIMT path and bf16 tcgen05 (MN-major operands, split-K) path vs float64;
    accumulation semantics (a second call doubles the result); the SIMT kernel on the same bf16 data as cross-check."""
    rng = np.random.Generator(np.random.PCG64(rows + N))
    x = torch.from_numpy(O.rng_normal(rng, (rows, K))).to(dtype)
    dy = torch.from_numpy(O.rng_normal(rng, (rows, N))).to(dtype)
    want_w = x.double().T @ dy.double()
    want_b = dy.double().sum(0)
    xd, dyd = x.to(cuda_dev), dy.to(cuda_dev)
    dw = torch.zeros((K, N), dtype=torch.float32, device=cuda_dev)
    db = torch.zeros((N,), dtype=torch.float32, device=cuda_dev)
    ops.linear_bwd_weight(xd, dyd, dw, db)
    tol = 1e-5 if dtype == torch.float32 else 1e-4          # bf16 inputs are exact products, fp32 accumulation
    assert rel_err(dw, want_w) < tol and rel_err(db, want_b) < tol
    ops.linear_bwd_weight(xd, dyd, dw, db)
    assert rel_err(dw, 2 * want_w) < tol and rel_err(db, 2 * want_b) < tol
    if dtype == torch.bfloat16:
        os.environ["EMRT_DW_SIMT"] = "1"
        try:
            dw2 = torch.zeros_like(dw)
            ops.linear_bwd_weight(xd, dyd, dw2, None)
        finally:
            del os.environ["EMRT_DW_SIMT"]
        assert rel_err(dw2, want_w) < tol


def test_module_window_gather_switch(cuda_dev):
    """MSDeformableAttention.window_gather = False takes the encoder self-attention through the L1-path gather (flat time,
    for models whose sampling offsets stray far from their reference points) instead of the window-staged kernels: same
    module output up to fp32 summation order, here with offsets wide enough (sigma 0.3) to leave the staged windows."""
    from emrt_b200 import synthetic
    shapes = synthetic.level_shapes(256)
    Lv = sum(h * w for h, w in shapes)
    g = torch.Generator(device=cuda_dev).manual_seed(11)
    src = torch.randn((2, Lv, 256), generator=g, device=cuda_dev).bfloat16()
    pos = torch.randn((1, Lv, 256), generator=g, device=cuda_dev).bfloat16()
    ref = emrt_b200.get_reference_points(shapes, device=cuda_dev)
    m = emrt_b200.MSDeformableAttention(256, 8, 3, 6).to(cuda_dev)
    with torch.no_grad():
        for name, arr in synthetic.msda_state(1234, offset_std=0.3).items():
            mod, leaf = name.split(".")
            getattr(getattr(m, mod), leaf).copy_(torch.from_numpy(arr))
    m.requires_grad_(False)
    assert m.window_gather is True
    with torch.no_grad():
        win = m(src, ref, src, shapes, query_pos=pos)
        m.window_gather = False
        l1 = m(src, ref, src, shapes, query_pos=pos)
    assert win.dtype == torch.bfloat16 and torch.isfinite(win.float()).all()
    d = (win.float() - l1.float()).norm() / win.float().norm()
    assert d.item() < 6e-3, d.item()
