"""Operator layer: one Python function per C-ABI entry point, taking torch CUDA tensors.

torch is used for device memory and streams only; all arithmetic happens inside libemrt_b200.so.
Every function raises on a non-CUDA tensor — there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib as L

_DT = {torch.float32: L.F32, torch.bfloat16: L.BF16, torch.float16: L.F16, torch.int32: L.I32, torch.uint8: L.U8}
_TD = {v: k for k, v in _DT.items()}


kernel_events = None   # set to a list to collect (name, dims, (start, end)) CUDA-event pairs per launch
_EVENT_POOL = []


def prewarm_events(n):
    """Create n timing events now (torch creates the cudaEvent at the first record, so each is recorded once, here, on the
    current stream): the C entries that record events in place then take them from this pool."""
    for _ in range(n):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        _EVENT_POOL.append(ev)


def _pooled_events(n):
    if len(_EVENT_POOL) < n:
        prewarm_events(n - len(_EVENT_POOL))
    out = _EVENT_POOL[-n:]
    del _EVENT_POOL[-n:]
    return out


class _Timed:
    """CUDA events around one launch on the launching stream when bench.py has set ``kernel_events`` to a list."""

    def __init__(self, name, dims):
        self.name, self.dims, self.ev = name, dims, None

    def __enter__(self):
        if kernel_events is not None:
            self.ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self.ev[0].record()
        return self

    def __exit__(self, *exc):
        if self.ev is not None:
            self.ev[1].record()
            kernel_events.append((self.name, self.dims, self.ev))
        return False


def _dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise L.EmrtError(f"unsupported dtype {t.dtype}")


def _ptr(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if not t.is_cuda:
        raise L.EmrtError("emrt_b200 ops need CUDA tensors (no CPU fallback)")
    if not t.is_contiguous():
        raise L.EmrtError("emrt_b200 ops need contiguous tensors")
    if t.device.index != torch.cuda.current_device():
        # the launch goes to the CURRENT device's current stream: a tensor living elsewhere would be a wrong-device launch
        raise L.EmrtError(f"tensor on {t.device} but the current device is cuda:{torch.cuda.current_device()}: "
                          "wrap the call in `with torch.cuda.device(tensor.device):`")
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def level_tables(shapes: Sequence[Tuple[int, int]]):
    hw, start, acc = [], [], 0
    for h, w in shapes:
        hw += [int(h), int(w)]
        start.append(acc)
        acc += int(h) * int(w)
    return L.i32_array(hw), L.i32_array(start), acc


def device_check():
    L.check(L.load().emrt_device_check())


def launch_count() -> int:
    return int(L.load().emrt_launch_count())


def reset_launch_count():
    L.load().emrt_reset_launch_count()


# ---- a3 ------------------------------------------------------------------------------------------------------
def msda_gather_fwd(value, loc, attn, shapes, ref=None, mode=L.LOC_NORMALIZED, out=None, win_center=None):
    """value [B,Lv,M,D]; loc [B,Lq,M,L,P,2]; attn [B,Lq,M,L,P]; ref [Bref,Lq,L,2] f32 (PIXEL_OFFSET) -> [B,Lq,M*D].
    win_center: optional host int32 array [M*L*2] from `L.i32_array` — the window-centre hint of
    emrt_msda_gather_fwd_hint (a locality hint for the window-staged kernel; results do not depend on it)."""
    lib = L.load()
    if mode & L.VALUE_HEAD_MAJOR:
        B, M, Lv, D = value.shape
    else:
        B, Lv, M, D = value.shape
    _, Lq, _, nL, P, _ = loc.shape
    hw, start, total = level_tables(shapes)
    if out is None:
        out = torch.empty((B, Lq, M * D), dtype=value.dtype, device=value.device)
    rbs = 0 if ref is None or ref.shape[0] == 1 else Lq * nL * 2
    ev = None
    if kernel_events is not None:      # bench.py: CUDA events around this launch, on the launching stream
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    L.check(lib.emrt_msda_gather_fwd_hint(_ptr(value), _ptr(loc), _ptr(attn), _ptr(ref), rbs, _ptr(out), B, Lq, Lv, M,
                                          D, nL, P, hw, start, _dt(value), _dt(loc), mode, win_center, _stream()))
    if ev is not None:
        ev[1].record()
        kernel_events.append(("msda_gather_fwd", (B, Lq, Lv, M, D, nL, P, value.element_size(), loc.element_size()), ev))
    return out


def msda_gather_bwd(grad_out, value, loc, attn, shapes, ref=None, mode=L.LOC_NORMALIZED, win_center=None):
    """-> grad_value f32 [B,Lv,M,D], grad_loc f32 [B,Lq,M,L,P,2], grad_attn f32 [B,Lq,M,L,P].  mode | QUERY_PIXEL_GRID
    selects the windowed backward (shared-memory fixed-point accumulation) on encoder geometry."""
    lib = L.load()
    B, Lv, M, D = value.shape
    _, Lq, _, nL, P, _ = loc.shape
    hw, start, total = level_tables(shapes)
    gv = torch.zeros((B, Lv, M, D), dtype=torch.float32, device=value.device)
    gl = torch.empty(tuple(loc.shape), dtype=torch.float32, device=value.device)
    ga = torch.empty(tuple(attn.shape), dtype=torch.float32, device=value.device)
    rbs = 0 if ref is None or ref.shape[0] == 1 else Lq * nL * 2
    L.check(lib.emrt_msda_gather_bwd_hint(_ptr(grad_out), _ptr(value), _ptr(loc), _ptr(attn), _ptr(ref), rbs, _ptr(gv),
                                          _ptr(gl), _ptr(ga), B, Lq, Lv, M, D, nL, P, hw, start, _dt(value), _dt(loc),
                                          mode, win_center, _stream()))
    return gv, gl, ga


# ---- nn.Linear -----------------------------------------------------------------------------------------------
def linear(x, w, bias=None, *, w_transposed=False, y_dtype=None, epilogue=L.EPI_NONE, row_scale=None, residual=None,
           ln_gamma=None, ln_beta=None, ln_eps=1e-5, qproj_group=0, impl=L.IMPL_AUTO, out=None, out2=None, hm_rows=0,
           hm_D=0, x2=None, x2_period=0, row_bias=None, row_bias_period=0, gn_branch=None, x_nchw=False):
    """y = epilogue((x + x2) @ W + bias).  x [..., K]; W [K,N] (Paddle layout) or [N,K] when w_transposed.
    x2 (optional, tcgen05 path): `cyclic_rows(addend, ...)` of a [x2_period, K] broadcast addend (with_pos_embed).
    gn_branch (EPI_RESIDUAL_LN only): dict(conv, skip, stats, gamma, beta, shapes, groups=32, eps=1e-5) — the encoder layer's
    conv branch GELU(GroupNorm_l(conv)) + skip added behind the LayerNorm inside the same epilogue."""
    lib = L.load()
    if x_nchw:
        # x [B, K, H, W] (or [B, K, hw]): the GEMM rows are the pixels, read channel-major by the tensor pipe (no transposed copy)
        assert x.dim() >= 3 and x.is_contiguous()
        K = x.shape[1]
        hw = x.numel() // (x.shape[0] * K)
        rows = x.shape[0] * hw
    else:
        K = x.shape[-1]
        rows = x.numel() // K
    N = w.shape[0] if w_transposed else w.shape[1]
    assert (w.shape[1] if w_transposed else w.shape[0]) == K, "weight/in-feature mismatch"
    ydt = y_dtype if y_dtype is not None else x.dtype
    if epilogue & L.EPI_MSDA_QPROJ:
        n_pts = N // 3
        if out is None:
            out = torch.empty((*x.shape[:-1], 2 * n_pts), dtype=ydt, device=x.device)
        if out2 is None:
            out2 = torch.empty((*x.shape[:-1], n_pts), dtype=ydt, device=x.device)
    elif out is None:
        out = torch.empty(((x.shape[0], hw, N) if x_nchw else (*x.shape[:-1], N)), dtype=ydt, device=x.device)
    a = L.LinearArgs()
    a.x, a.w, a.bias, a.y = _ptr(x), _ptr(w), _ptr(bias), _ptr(out)
    a.rows, a.K, a.N = rows, K, N
    a.x_dtype, a.w_dtype, a.y_dtype, a.w_transposed = _dt(x), _dt(w), _DT[ydt], int(bool(w_transposed))
    a.epilogue = int(epilogue)
    a.row_scale = _ptr(row_scale)
    a.residual, a.ln_gamma, a.ln_beta, a.ln_eps = _ptr(residual), _ptr(ln_gamma), _ptr(ln_beta), float(ln_eps)
    a.y2 = _ptr(out2)
    a.qproj_group = int(qproj_group)
    a.impl = int(impl)
    a.hm_rows, a.hm_D = int(hm_rows), int(hm_D)
    a.x2, a.x2_period = _ptr(x2), int(x2_period)
    a.row_bias, a.row_bias_period = _ptr(row_bias), int(row_bias_period)
    a.x_nchw_hw = int(hw) if x_nchw else 0
    gnb = None
    if gn_branch is not None:
        gnb = L.GnBranch()
        gnb.conv, gnb.skip, gnb.stats = _ptr(gn_branch["conv"]), _ptr(gn_branch["skip"]), _ptr(gn_branch["stats"])
        gnb.gamma, gnb.beta = _ptr(gn_branch["gamma"]), _ptr(gn_branch["beta"])
        shapes = gn_branch["shapes"]
        gnb.L, gnb.groups, gnb.eps = len(shapes), int(gn_branch.get("groups", 32)), float(gn_branch.get("eps", 1e-5))
        gnb.Lv = sum(int(h) * int(w) for h, w in shapes)
        for i, (h, w) in enumerate(shapes):
            gnb.shapes_hw[2 * i], gnb.shapes_hw[2 * i + 1] = int(h), int(w)
        a.gn = C.pointer(gnb)
    if row_bias is not None:
        assert row_bias.dtype == torch.float16 and row_bias.shape[-1] == N and row_bias.numel() // N >= row_bias_period + 127
    if x2 is not None:
        assert x2.dtype == torch.bfloat16 and x2.shape[-1] == K and x2.numel() // K >= x2_period + 127
    with _Timed("linear_ln" if (epilogue & L.EPI_RESIDUAL_LN) else "linear", (rows, K, N, x.element_size(), out.element_size())):
        L.check(lib.emrt_linear_fwd(C.byref(a), _stream()))
    return (out, out2) if (epilogue & L.EPI_MSDA_QPROJ) else out


def _gn_struct(gn_branch):
    gnb = L.GnBranch()
    gnb.conv, gnb.skip, gnb.stats = _ptr(gn_branch["conv"]), _ptr(gn_branch["skip"]), _ptr(gn_branch["stats"])
    gnb.gamma, gnb.beta = _ptr(gn_branch["gamma"]), _ptr(gn_branch["beta"])
    shapes = gn_branch["shapes"]
    gnb.L, gnb.groups, gnb.eps = len(shapes), int(gn_branch.get("groups", 32)), float(gn_branch.get("eps", 1e-5))
    gnb.Lv = sum(int(h) * int(w) for h, w in shapes)
    for i, (h, w) in enumerate(shapes):
        gnb.shapes_hw[2 * i], gnb.shapes_hw[2 * i + 1] = int(h), int(w)
    return gnb


def ffn_fused(x, w1, b1, w2, b2, ln_gamma, ln_beta, *, ln_eps=1e-5, gn_branch=None, out=None):
    """y = LayerNorm(x + linear2(relu(linear1(x)))) * ln_gamma + ln_beta (+ conv branch): forward_ffn (t_e_d.py:157-160)
    and the layer's final add (:203) in one kernel; the hidden tensor never reaches HBM.  x bf16 [..., 256]; w1 / w2 the
    packed bf16 [d_ff, 256] / [256, d_ff] weights; gn_branch as in `linear`."""
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and w1.dtype == torch.bfloat16 and w2.dtype == torch.bfloat16
    d_model = x.shape[-1]
    d_ff = w1.shape[0]
    assert w1.shape == (d_ff, d_model) and w2.shape == (d_model, d_ff)
    if out is None:
        out = torch.empty_like(x)
    a = L.FfnArgs()
    a.x, a.w1, a.b1, a.w2, a.b2 = _ptr(x), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2)
    a.ln_gamma, a.ln_beta, a.ln_eps = _ptr(ln_gamma), _ptr(ln_beta), float(ln_eps)
    a.y, a.rows, a.d_model, a.d_ff = _ptr(out), x.numel() // d_model, d_model, d_ff
    gnb = None
    if gn_branch is not None:
        gnb = _gn_struct(gn_branch)
        a.gn = C.pointer(gnb)
    with _Timed("ffn_fused", (a.rows, d_model, d_ff, x.element_size())):
        L.check(L.load().emrt_ffn_fused_fwd(C.byref(a), _stream()))
    return out


def cyclic_rows(addend, extra=127, dtype=torch.bfloat16):
    """[period, K] -> `dtype` [period + extra, K]: the rows continued cyclically (the x2 / row_bias operand of `linear`:
    any 128-row tile of a [B * period, .] activation then reads its addend rows as one contiguous box)."""
    period = addend.shape[0]
    idx = torch.arange(period + extra, device=addend.device) % period
    return addend.to(dtype).index_select(0, idx).contiguous()


_cyc_cache = {}


def cyclic_rows_cached(addend):
    """`cyclic_rows` remembered for as long as that very tensor object is alive and unmodified."""
    import weakref
    hit = _cyc_cache.get(id(addend))
    if hit is not None and hit[0]() is addend and hit[1] == addend._version:
        return hit[2]
    val = cyclic_rows(addend.reshape(-1, addend.shape[-1]))
    if len(_cyc_cache) > 64:
        _cyc_cache.clear()
    _cyc_cache[id(addend)] = (weakref.ref(addend), addend._version, val)
    return val


def pack_weight(src, dst, dst_row0=0):
    """dst[dst_row0 + n, k] = bf16(src[k, n]); src is a Paddle-layout [K,N] weight."""
    K, N = src.shape
    assert dst.dtype == torch.bfloat16 and dst.shape[1] == K and dst.shape[0] >= dst_row0 + N
    L.check(L.load().emrt_pack_weight(_ptr(src), _dt(src), _ptr(dst), K, N, int(dst_row0), _stream()))
    return dst


def msda_softmax_loc(off_raw, logit_raw, shapes, M, P, ref=None, out_dtype=torch.float32, mode=L.LOC_NORMALIZED):
    """off_raw f32 [B,Lq,M*L*P*2] (may be a column slice view with row stride), logit_raw f32 [B,Lq,M*L*P]."""
    lib = L.load()
    B, Lq = off_raw.shape[:2]
    nL = len(shapes)
    hw, _, _ = level_tables(shapes)
    for t in (off_raw, logit_raw):
        if not t.is_cuda or t.dtype != torch.float32 or t.stride(-1) != 1:
            raise L.EmrtError("softmax_loc needs f32 CUDA inputs with unit inner stride")
    loc = torch.empty((B, Lq, M, nL, P, 2), dtype=out_dtype, device=off_raw.device)
    attn = torch.empty((B, Lq, M, nL, P), dtype=out_dtype, device=off_raw.device)
    rbs = 0 if ref is None or ref.shape[0] == 1 else Lq * nL * 2
    L.check(lib.emrt_msda_softmax_loc(C.c_void_p(off_raw.data_ptr()), off_raw.stride(-2),
                                      C.c_void_p(logit_raw.data_ptr()), logit_raw.stride(-2), _ptr(ref), rbs,
                                      _ptr(loc), _ptr(attn), B, Lq, M, nL, P, hw, _DT[out_dtype], mode, _stream()))
    return loc, attn


# ---- backward pieces ---------------------------------------------------------------------------------------------
def linear_bwd_weight(x, dy, dw, db=None):
    """dw [K,N] f32 += x[rows,K]^T dy[rows,N]; db [N] f32 += column sums of dy.  Accumulates into dw / db."""
    K, N = dw.shape
    rows = x.numel() // K
    assert dy.numel() == rows * N and dw.dtype == torch.float32
    L.check(L.load().emrt_linear_bwd_weight(_ptr(x), _ptr(dy), _ptr(dw), _ptr(db), rows, K, N, _dt(x), _dt(dy), _stream()))
    return dw, db


def msda_qproj_bwd(grad_loc, grad_attn, attn, shapes, M, P, out_dtype=torch.float32, mode=L.LOC_NORMALIZED):
    """-> dq [..., 3*M*L*P] = [offset grads | logit grads] (softmax + location backward)."""
    nL = len(shapes)
    lead = attn.shape[:-3]
    rows = attn.numel() // (M * nL * P)
    hw, _, _ = level_tables(shapes)
    dq = torch.empty((*lead, 3 * M * nL * P), dtype=out_dtype, device=attn.device)
    L.check(L.load().emrt_msda_qproj_bwd(_ptr(grad_loc), _ptr(grad_attn), _ptr(attn), _ptr(dq), rows, M, nL, P, hw,
                                         _dt(attn), _DT[out_dtype], mode, _stream()))
    return dq


def msda_ref_bwd(grad_loc, shapes, ref_batches, mode=L.LOC_NORMALIZED):
    """grad_loc f32 [B,Lq,M,L,P,2] -> grad of the reference points f32 [ref_batches,Lq,L,2] (t_e_d.py:98-102)."""
    B, Lq, M, nL, P, _ = grad_loc.shape
    hw, _, _ = level_tables(shapes)
    out = torch.empty((ref_batches, Lq, nL, 2), dtype=torch.float32, device=grad_loc.device)
    L.check(L.load().emrt_msda_ref_bwd(_ptr(grad_loc), _ptr(out), B, ref_batches, Lq, M, nL, P, hw, mode & 1, _stream()))
    return out


def scale_rows_cast(src, row_scale, out_dtype):
    """dst[r, :] = src[r, :] * row_scale[r] (row_scale may be None), fp32 -> out_dtype."""
    cols = src.shape[-1]
    rows = src.numel() // cols
    dst = torch.empty(src.shape, dtype=out_dtype, device=src.device)
    L.check(L.load().emrt_scale_rows_cast(_ptr(src), _ptr(row_scale), _ptr(dst), rows, cols, _DT[out_dtype], _stream()))
    return dst


def add_layernorm(x, residual, gamma, beta, eps=1e-5, out=None):
    N = x.shape[-1]
    rows = x.numel() // N
    if out is None:
        out = torch.empty_like(x)
    L.check(L.load().emrt_add_layernorm(_ptr(x), _ptr(residual), _ptr(gamma), _ptr(beta), _ptr(out), rows, N,
                                        float(eps), _dt(x), _stream()))
    return out


def residual_layernorm(x, residual, gamma, beta, post_add=None, eps=1e-5, out=None):
    """y = LayerNorm(x + residual) * gamma + beta (+ post_add) over the last dim (t_e_d.py:199-200,159-160,203)."""
    N = x.shape[-1]
    rows = x.numel() // N
    if out is None:
        out = torch.empty_like(x)
    L.check(L.load().emrt_residual_layernorm(_ptr(x), _ptr(residual), _ptr(gamma), _ptr(beta), _ptr(post_add), _ptr(out),
                                             rows, N, float(eps), _dt(x), _stream()))
    return out


def pack_conv3x3_weights(weights, dtype):
    """[Cout,Cin,3,3] fp32 per level (Paddle Conv2D layout) -> [L, 9, Cout, Cin] `dtype` conv operand."""
    C = weights[0].shape[0]
    dst = torch.empty((len(weights), 9, C, C), dtype=dtype, device=weights[0].device)
    for l, w in enumerate(weights):
        assert tuple(w.shape) == (C, C, 3, 3)
        L.check(L.load().emrt_pack_conv3x3_weight(_ptr(w.detach().float().contiguous()), _ptr(dst), C, l, _DT[dtype], _stream()))
    return dst


def conv3x3_tokens(x, w_packed, shapes, impl=L.IMPL_AUTO, out=None):
    """Per-level 3x3 conv (pad 1, no bias) on tokens x [B, Lv, C]; w_packed from pack_conv3x3_weights."""
    B, Lv, C = x.shape
    hw, _, total = level_tables(shapes)
    assert total == Lv
    if out is None:
        out = torch.empty_like(x)
    with _Timed("conv3x3", (B * Lv, C, x.element_size())):
        L.check(L.load().emrt_conv3x3_tokens_fwd(_ptr(x), _ptr(w_packed), _ptr(out), B, Lv, C, len(shapes), hw, _dt(x),
                                                 _dt(w_packed), int(impl), _stream()))
    return out


def conv3x3_tokens_stats(x, w_packed, shapes, groups=32, out=None, stats=None, max_ctas=0):
    """conv3x3_tokens on the tcgen05 path + the GroupNorm(groups) statistics of its stored output from the same kernel's
    epilogue -> (y, stats workspace: its first 2 * B * L * groups floats are the sums [B, L, groups, 2], as groupnorm_stats').
    Raises EmrtError(UNSUPPORTED) for shapes the tcgen05 conv does not tile.  max_ctas > 0: on at most that many SMs (for a
    caller that runs it on a second stream beside another kernel; `out` / `stats` = conv3x3_stats_buffers(...) allocated on
    the consumer's stream)."""
    B, Lv, C = x.shape
    hw, _, total = level_tables(shapes)
    assert total == Lv and x.dtype == torch.bfloat16 and w_packed.dtype == torch.bfloat16
    if out is None:
        out = torch.empty_like(x)
    lib = L.load()
    st = stats if stats is not None else conv3x3_stats_buffers(x, shapes)[1]
    with _Timed("conv3x3", (B * Lv, C, x.element_size())):
        L.check(lib.emrt_conv3x3_tokens_stats_part_fwd(_ptr(x), _ptr(w_packed), _ptr(out), _ptr(st), B, Lv, C, len(shapes), hw,
                                                       int(groups), int(max_ctas), _stream()))
    return out, st


def conv3x3_stats_buffers(x, shapes):
    """(output, statistics workspace) of conv3x3_tokens_stats for an input like x, allocated on the current stream."""
    B, Lv, C = x.shape
    n = int(L.load().emrt_conv3x3_stats_workspace_floats(B, Lv, len(shapes)))
    return torch.empty_like(x), torch.empty((n,), dtype=torch.float32, device=x.device)


def groupnorm_gelu_residual(conv, x, gamma, beta, shapes, groups=32, eps=1e-5, out=None, return_stats=False):
    """GELU(GroupNorm_l(conv)) + x per level; gamma / beta f32 [L, C].  return_stats: also the statistics workspace
    (its first 2 * B * L * groups floats are the (sum, sum of squares) the backward needs)."""
    B, Lv, C = x.shape
    hw, _, _ = level_tables(shapes)
    if out is None:
        out = torch.empty_like(x)
    ws = torch.empty((int(L.load().emrt_groupnorm_workspace_floats(B, len(shapes), groups)),), dtype=torch.float32, device=x.device)
    L.check(L.load().emrt_groupnorm_gelu_residual(_ptr(conv), _ptr(x), _ptr(gamma), _ptr(beta), _ptr(out), _ptr(ws), B, Lv,
                                                  C, len(shapes), groups, float(eps), hw, _dt(x), _stream()))
    return (out, ws) if return_stats else out


def groupnorm_stats(x, shapes, groups=32):
    """(sum, sum of squares) per (batch, level, group) of x [B, Lv, C] -> the statistics workspace whose first
    2 * B * L * groups floats are the sums [B, L, groups, 2] (deterministic fixed-order reduction)."""
    B, Lv, C = x.shape
    hw, _, _ = level_tables(shapes)
    st = torch.empty((int(L.load().emrt_groupnorm_workspace_floats(B, len(shapes), groups)),), dtype=torch.float32, device=x.device)
    L.check(L.load().emrt_groupnorm_stats(_ptr(x), _ptr(st), B, Lv, C, len(shapes), groups, hw, _dt(x), _stream()))
    return st       # the sums [B, L, groups, 2] are its first 2 * B * L * groups floats


def residual_layernorm_gn(x, residual, ln_gamma, ln_beta, conv, skip, gn_stats, gn_gamma, gn_beta, shapes, groups=32,
                          ln_eps=1e-5, gn_eps=1e-5, out=None):
    """y = LayerNorm(x + residual) * ln_gamma + ln_beta + GELU(GroupNorm_l(conv)) + skip (t_e_d.py:159-160,187-189,203):
    the conv branch is evaluated inside the LayerNorm pass from the conv output and groupnorm_stats' sums."""
    B, Lv, C = x.shape
    hw, _, _ = level_tables(shapes)
    if out is None:
        out = torch.empty_like(x)
    L.check(L.load().emrt_residual_layernorm_gn(_ptr(x), _ptr(residual), _ptr(ln_gamma), _ptr(ln_beta), _ptr(conv), _ptr(skip),
                                                _ptr(gn_stats), _ptr(gn_gamma), _ptr(gn_beta), _ptr(out), B, Lv, C,
                                                len(shapes), groups, float(ln_eps), float(gn_eps), hw, _dt(x), _stream()))
    return out


def nchw_to_tokens(x):
    """[B, C, *spatial] -> tokens [B, prod(spatial), C]."""
    B, C_ = x.shape[:2]
    P = x.numel() // (B * C_)
    y = torch.empty((B, P, C_), dtype=x.dtype, device=x.device)
    L.check(L.load().emrt_nchw_to_tokens(_ptr(x.contiguous()), _ptr(y), B, C_, P, _dt(x), _stream()))
    return y


def groupnorm_tokens_into(x, gamma, beta, out, token_offset, groups=32, eps=1e-5, return_stats=False):
    """GroupNorm of one level's tokens x [B, P, C] written into out[:, token_offset:token_offset+P, :] (out [B, Lv, C])."""
    B, P, C_ = x.shape
    assert out.is_contiguous() and out.shape[0] == B and out.shape[2] == C_ and out.dtype == x.dtype
    ws = torch.empty((int(L.load().emrt_groupnorm_workspace_floats(B, 1, groups)),), dtype=torch.float32, device=x.device)
    dst = C.c_void_p(out.data_ptr() + token_offset * C_ * out.element_size())
    L.check(L.load().emrt_groupnorm_tokens(_ptr(x), _ptr(gamma), _ptr(beta), dst, out.shape[1] * C_, _ptr(ws), B, P, C_,
                                           groups, float(eps), _dt(x), _stream()))
    return (out, ws) if return_stats else out


def mha_small(q, k, v, num_heads, scale):
    """softmax(scale * q k^T) v per head on projected q/k/v [B, L, C] (last-dim-contiguous views allowed) -> [B, Lq, C]."""
    B, Lq, C_ = q.shape
    Lk = k.shape[1]
    for t in (q, k, v):
        if not t.is_cuda or t.stride(-1) != 1 or t.stride(0) != t.shape[1] * t.stride(1):
            raise L.EmrtError("mha_small needs CUDA tensors with unit inner stride and a uniform row stride")
    out = torch.empty((B, Lq, C_), dtype=q.dtype, device=q.device)
    p = lambda t: C.c_void_p(t.data_ptr())
    L.check(L.load().emrt_mha_small(p(q), q.stride(1), p(k), k.stride(1), p(v), v.stride(1), _ptr(out), B, Lq, Lk,
                                    num_heads, C_ // num_heads, float(scale), _dt(q), _stream()))
    return out


def add_bcast(a, b, out=None):
    """out = a + b with b broadcast over leading dims (b.numel() divides a.numel()): with_pos_embed."""
    if out is None:
        out = torch.empty_like(a)
    L.check(L.load().emrt_add_bcast(_ptr(a), _ptr(b), _ptr(out), a.numel(), b.numel(), _dt(a), _stream()))
    return out


# ---- head tail -----------------------------------------------------------------------------------------------
def upsample2x(x):
    n, nc, h, w = x.shape
    out = torch.empty((n, nc, 2 * h, 2 * w), dtype=torch.float32, device=x.device)
    L.check(L.load().emrt_upsample2x(_ptr(x), _ptr(out), n, nc, h, w, _dt(x), _stream()))
    return out


def window_accumulate(win_logits, canvas, count, win_img, win_y0, win_x0):
    n_win, nc, hc, wc = win_logits.shape
    n_img, _, H, W = canvas.shape
    assert win_logits.dtype == torch.float32 and canvas.dtype == torch.float32
    L.check(L.load().emrt_window_accumulate(_ptr(win_logits), _ptr(canvas), _ptr(count), n_win, n_img, nc, hc, wc, H,
                                            W, _ptr(win_img), _ptr(win_y0), _ptr(win_x0), _stream()))
    return canvas, count


def finalize_argmax(canvas, count=None, out_hw=None, label_dtype=torch.int32, want_probs=False, want_logits=False):
    n_img, nc, H, W = canvas.shape
    Ho, Wo = (H, W) if out_hw is None else (int(out_hw[0]), int(out_hw[1]))
    labels = torch.empty((n_img, 1, Ho, Wo), dtype=label_dtype, device=canvas.device)
    probs = torch.empty((n_img, nc, Ho, Wo), dtype=torch.float32, device=canvas.device) if want_probs else None
    logits = torch.empty((n_img, nc, H, W), dtype=torch.float32, device=canvas.device) if want_logits else None
    L.check(L.load().emrt_finalize_argmax(_ptr(canvas), _ptr(count), _ptr(labels), _DT[label_dtype], _ptr(probs),
                                          _ptr(logits), n_img, nc, H, W, Ho, Wo, _stream()))
    return labels, probs, logits


def stitch_argmax_fused(half_logits, win_img, win_y0, win_x0, n_img, H, W, label_dtype=torch.int32,
                        want_logits=False, labels=None):
    n_win, nc, hh, hw = half_logits.shape
    if labels is None:
        labels = torch.empty((n_img, 1, H, W), dtype=label_dtype, device=half_logits.device)
    logits = torch.empty((n_img, nc, H, W), dtype=torch.float32, device=half_logits.device) if want_logits else None
    L.check(L.load().emrt_stitch_argmax_fused(_ptr(half_logits), _dt(half_logits), _ptr(labels), _dt(labels),
                                              _ptr(logits), n_win, n_img, nc, 2 * hh, 2 * hw, H, W, _ptr(win_img),
                                              _ptr(win_y0), _ptr(win_x0), _stream()))
    return labels, logits


def stitch_argmax_eval(half_logits, win_img, win_y0, win_x0, n_img, H, W, gt=None, ignore_index=255, palette=None,
                       label_dtype=torch.int32, labels=None):
    """stitch_argmax_fused with the evaluation fused behind the argmax.
    gt: ground-truth labels [n_img, (1,) H, W] (any integer dtype; stored as int32 / uint8) -> areas int64 [n_img, 3, nc]
    (metrics.calculate_area per image); palette: uint8 [nc, 3] -> colour image uint8 [n_img, H, W, 3] (predict.py:171-174).
    Returns (labels, areas | None, color | None)."""
    n_win, nc, hh, hw = half_logits.shape
    dev = half_logits.device
    if labels is None:
        labels = torch.empty((n_img, 1, H, W), dtype=label_dtype, device=dev)
    areas = color = None
    if gt is not None:
        gt = gt.reshape(n_img, H, W)
        if gt.dtype not in (torch.int32, torch.uint8):
            gt = gt.to(torch.int32)
        gt = gt.contiguous()
        areas = torch.zeros((n_img, 3, nc), dtype=torch.int64, device=dev)
    if palette is not None:
        palette = palette.to(device=dev, dtype=torch.uint8).reshape(nc, 3).contiguous()
        color = torch.empty((n_img, H, W, 3), dtype=torch.uint8, device=dev)
    L.check(L.load().emrt_stitch_argmax_eval(_ptr(half_logits), _dt(half_logits), _ptr(labels), _dt(labels), n_win, n_img, nc,
                                             2 * hh, 2 * hw, H, W, _ptr(win_img), _ptr(win_y0), _ptr(win_x0), _ptr(gt),
                                             _dt(gt) if gt is not None else 0, int(ignore_index), _ptr(areas), _ptr(palette),
                                             _ptr(color), _stream()))
    return labels, areas, color


def calculate_area(pred, label, num_classes, ignore_index=255):
    """-> int64 [3, num_classes] = (intersect, pred, label) areas (src/utils/metrics.py:20-69)."""
    pred = pred.reshape(-1).to(torch.int32).contiguous()
    label = label.reshape(-1).to(torch.int32).contiguous()
    if pred.shape != label.shape:
        raise ValueError("Shape of `pred` and `label should be equal")
    areas = torch.zeros((3, num_classes), dtype=torch.int64, device=pred.device)
    L.check(L.load().emrt_calculate_area(_ptr(pred), _ptr(label), pred.numel(), num_classes, int(ignore_index),
                                         _ptr(areas), _stream()))
    return areas


# ---- backward of the encoder / decoder glue (cfg 4) -----------------------------------------------------------------
def layernorm_bwd(a, b, gamma, dy, dgamma, dbeta, eps=1e-5):
    """Backward of y = LayerNorm(a + b) * gamma + beta: returns dz (the gradient of a and of b); dgamma / dbeta f32 [N] are
    accumulated in place."""
    N = a.shape[-1]
    rows = a.numel() // N
    lib = L.load()
    ws = torch.empty((int(lib.emrt_layernorm_bwd_workspace_floats(rows, N)) + 2 * N,), dtype=torch.float32, device=a.device)
    dz = torch.empty_like(a)
    L.check(lib.emrt_layernorm_bwd(_ptr(a), _ptr(b), _ptr(gamma), _ptr(dy), _ptr(dz), _ptr(dgamma), _ptr(dbeta), _ptr(ws),
                                   rows, N, float(eps), _dt(a), _stream()))
    return dz


def groupnorm_bwd(x, dy, stats, gamma, beta, dgamma, dbeta, shapes, groups=32, eps=1e-5, gelu=False):
    """Backward of GroupNorm_l (+ GELU) on tokens x [B, Lv, C]: returns dx; dgamma / dbeta f32 [L, C] accumulated in place.
    stats: groupnorm_stats(x) (or the workspace a forward returned with return_stats=True)."""
    B, Lv, C = x.shape
    hw, _, total = level_tables(shapes)
    assert total == Lv
    lib = L.load()
    ws = torch.empty((int(lib.emrt_groupnorm_bwd_workspace_floats(B, len(shapes), C, groups)),), dtype=torch.float32, device=x.device)
    dx = torch.empty_like(x)
    L.check(lib.emrt_groupnorm_bwd(_ptr(x), _ptr(dy), _ptr(stats), _ptr(gamma), _ptr(beta), _ptr(dx), _ptr(dgamma), _ptr(dbeta),
                                   _ptr(ws), B, Lv, C, len(shapes), groups, float(eps), hw, int(bool(gelu)), _dt(x), _stream()))
    return dx


def relu_bwd(dy, y, out=None):
    if out is None:
        out = torch.empty_like(dy)
    L.check(L.load().emrt_relu_bwd(_ptr(dy), _ptr(y), _ptr(out), dy.numel(), _dt(dy), _stream()))
    return out


def batch_sum(x):
    """x [B, ...] -> f32 [...]: sum over the leading dimension."""
    B = x.shape[0]
    out = torch.empty(x.shape[1:], dtype=torch.float32, device=x.device)
    L.check(L.load().emrt_batch_sum(_ptr(x), _ptr(out), B, x.numel() // B, _dt(x), _stream()))
    return out


def column_sum(x, out=None):
    """x [rows, N] -> f32 [N] (accumulated into `out` when given)."""
    N = x.shape[-1]
    if out is None:
        out = torch.zeros((N,), dtype=torch.float32, device=x.device)
    L.check(L.load().emrt_column_sum(_ptr(x), _ptr(out), x.numel() // N, N, _dt(x), _stream()))
    return out


def sigmoid(x):
    y = torch.empty_like(x)
    L.check(L.load().emrt_sigmoid_fwd(_ptr(x), _ptr(y), x.numel(), _stream()))
    return y


def sigmoid_bwd(dy, y):
    dx = torch.empty_like(y)
    L.check(L.load().emrt_sigmoid_bwd(_ptr(dy), _ptr(y), _ptr(dx), y.numel(), _stream()))
    return dx


def mha_small_bwd(q, k, v, d_out, num_heads, scale):
    """Backward of mha_small: -> (dqk [B, L, 2C] = [dq | dk] when Lq == Lk else (dq, dk), dv)."""
    B, Lq, C_ = q.shape
    Lk = k.shape[1]
    dev = q.device
    if Lq == Lk:
        dqk = torch.empty((B, Lq, 2 * C_), dtype=q.dtype, device=dev)
        dq, dk = dqk[..., :C_], dqk[..., C_:]
    else:
        dqk = None
        dq = torch.empty((B, Lq, C_), dtype=q.dtype, device=dev)
        dk = torch.empty((B, Lk, C_), dtype=q.dtype, device=dev)
    dv = torch.empty((B, Lk, C_), dtype=q.dtype, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    L.check(L.load().emrt_mha_small_bwd(p(q), q.stride(1), p(k), k.stride(1), p(v), v.stride(1), _ptr(d_out.contiguous()),
                                        p(dq), dq.stride(1), p(dk), dk.stride(1), p(dv), dv.stride(1), B, Lq, Lk, num_heads,
                                        C_ // num_heads, float(scale), _dt(q), _stream()))
    return (dqk if dqk is not None else (dq, dk)), dv


def conv3x3_tokens_bwd_weight(x, dy, shapes, impl=L.IMPL_AUTO):
    """-> dw f32 [L, Cout, Cin, 3, 3]: the weight gradient of conv3x3_tokens per level (Paddle Conv2D layout)."""
    B, Lv, C_ = x.shape
    hw, _, total = level_tables(shapes)
    assert total == Lv
    nL = len(shapes)
    dw = torch.zeros((nL, C_, C_, 3, 3), dtype=torch.float32, device=x.device)
    ws = torch.empty((nL * 9 * C_ * C_,), dtype=torch.float32, device=x.device)
    L.check(L.load().emrt_conv3x3_tokens_bwd_weight(_ptr(x), _ptr(dy), _ptr(dw), _ptr(ws), B, Lv, C_, nL, hw, _dt(x), int(impl),
                                                    _stream()))
    return dw


def tokens_to_nchw(y, spatial):
    """[B, P, C] -> [B, C, *spatial]: the inverse of nchw_to_tokens (the same batched transpose with the roles swapped)."""
    B, P, C_ = y.shape
    x = torch.empty((B, C_, P), dtype=y.dtype, device=y.device)
    L.check(L.load().emrt_nchw_to_tokens(_ptr(y.contiguous()), _ptr(x), B, P, C_, _dt(y), _stream()))
    return x.view(B, C_, *spatial)


# ---- a2 in one call ---------------------------------------------------------------------------------------------------
class _Keep:
    """the ctypes argument struct of a fused MSDA call together with every tensor it points to"""

    def __init__(self, args, tensors):
        self.args, self.tensors = args, tensors


def msda_fused_fwd(query, value, ref, shapes, M, P, weights, *, mask=None, query_pos=None, query_pos_rows=0, row_bias=None,
                   residual_norm=None, pixel_grid=False, win_center=None, keep_pixel_major=False, out=None,
                   gather_start_event=None):
    """emrt_msda_fused_fwd: MSDeformableAttention.forward (t_e_d.py:65-107) in ONE C call.  query [B,Lq,C], value [B,Lv,C]
    (fp32 -> parity path, bf16 -> B200 path), ref f32 [1|B,Lq,L,2].  `weights`: fp32 path dict(w_value, b_value, w_offsets,
    b_offsets, w_attn, b_attn, w_out, b_out) in Paddle layout; bf16 path dict(wv, wq, wo, bv, bq, bo) packed (+ the fp32-path
    keys when a backward will follow).  Returns (out, keep): `keep` holds the argument struct and the workspace with the
    intermediates (what msda_fused_bwd needs)."""
    lib = L.load()
    B, Lq, C_ = query.shape
    Lv = value.shape[1]
    nL = len(shapes)
    dt = _dt(query)
    a = L.MsdaArgs()
    keep = [query, value, ref, mask, query_pos, row_bias, weights]
    a.query, a.value, a.ref, a.ref_batches = _ptr(query), _ptr(value), _ptr(ref), int(ref.shape[0])
    a.value_mask = _ptr(mask)
    for k_c, k_py in (("w_value", "w_value"), ("b_value", "b_value"), ("w_offsets", "w_offsets"), ("b_offsets", "b_offsets"),
                      ("w_attn", "w_attn"), ("b_attn", "b_attn"), ("w_out", "w_out"), ("b_out", "b_out"),
                      ("wv_packed", "wv"), ("wq_packed", "wq"), ("wo_packed", "wo"), ("b_query", "bq")):
        if weights.get(k_py) is not None:
            setattr(a, k_c, _ptr(weights[k_py]))
    if dt == L.BF16:
        a.b_value, a.b_out = _ptr(weights["bv"]), _ptr(weights["bo"])
    tp = M * nL * P
    scratch = None
    if query_pos is not None:
        a.query_pos, a.query_pos_rows = _ptr(query_pos), int(query_pos_rows)
        if dt == L.F32:
            scratch = torch.empty_like(query)
            a.query_eff = _ptr(scratch)
    if dt == L.BF16 and not (tp == 144 and nL * P == 18):
        scratch = torch.empty((B, Lq, 3 * tp), dtype=torch.float32, device=query.device)
    a.query_scratch = _ptr(scratch)
    a.row_bias = _ptr(row_bias)
    if residual_norm is not None:
        res, gamma, beta = residual_norm
        a.residual, a.ln_gamma, a.ln_beta, a.ln_eps = _ptr(res), _ptr(gamma), _ptr(beta), 1e-5
        keep += [res, gamma, beta]
    if out is None:
        out = torch.empty((B, Lq, C_), dtype=query.dtype, device=query.device)
    ws = torch.empty((int(lib.emrt_msda_fused_workspace_bytes(B, Lq, Lv, C_, M, nL, P, dt)),), dtype=torch.uint8, device=query.device)
    a.out, a.workspace = _ptr(out), _ptr(ws)
    a.B, a.Lq, a.Lv, a.C, a.M, a.L, a.P = B, Lq, Lv, C_, M, nL, P
    for i, (h, w) in enumerate(shapes):
        a.shapes_hw[2 * i], a.shapes_hw[2 * i + 1] = int(h), int(w)
    a.dtype, a.flags, a.keep_pixel_major = dt, (L.QUERY_PIXEL_GRID if pixel_grid else 0), int(bool(keep_pixel_major))
    if win_center is not None and pixel_grid:
        a.window_center = C.cast(win_center, C.POINTER(C.c_int32))
        keep.append(win_center)
    if gather_start_event is not None:      # a torch.cuda.Event the C entry records right before the gather launch
        gather_start_event.record()         # (recorded once here so that torch has created it; re-recorded in place)
        a.gather_start_event = gather_start_event.cuda_event
    evs = None
    if kernel_events is not None and dt == L.BF16:
        # bench.py: CUDA events around each sub-launch, recorded by the C entry on the launching stream.  Each event is
        # recorded once here so that torch has created it (and knows it as recorded); the C side records it again in place.
        # Five events for the four sub-launches (the end of one is the start of the next), taken from the pool bench.py filled
        # before its timed region when it did: fewer records inside it.
        five = _pooled_events(5)
        evs = [five[0], five[1], five[1], five[2], five[2], five[3], five[3], five[4]]
        for i, ev in enumerate(evs):
            a.timing_events[i] = ev.cuda_event
    L.check(lib.emrt_msda_fused_fwd(C.byref(a), _stream()))
    if evs is not None:
        es = query.element_size()
        kernel_events.append(("linear", (B * Lv, C_, C_, es, es), (evs[0], evs[1])))
        kernel_events.append(("linear", (B * Lq, C_, 3 * tp, es, 2), (evs[2], evs[3])))
        kernel_events.append(("msda_gather_fwd", (B, Lq, Lv, M, C_ // M, nL, P, es, 2), (evs[4], evs[5])))
        # the output projection carries residual + LayerNorm in its epilogue when residual_norm is given: its own row (it
        # also reads the residual), not one more sample of the plain K = N = C projection
        kernel_events.append(("linear_ln" if residual_norm is not None else "linear", (B * Lq, C_, C_, es, es), (evs[6], evs[7])))
    return out, _Keep(a, keep + [scratch, ws, out, evs])


def msda_fused_bwd(keep, d_out, wq_cat, w_value_cast=None, w_out_cast=None, want_ref_grad=False):
    """emrt_msda_fused_bwd on the argument struct a msda_fused_fwd(..., keep_pixel_major=True) returned.
    -> (d_query, d_value, d_ref | None, dw_query [C,3MLP], db_query, dw_value, db_value, dw_out, db_out), weights grads f32."""
    lib = L.load()
    a = keep.args
    B, Lq, Lv, C_, M, nL, P = a.B, a.Lq, a.Lv, a.C, a.M, a.L, a.P
    tp = M * nL * P
    dev = d_out.device
    adt = d_out.dtype
    f32 = dict(dtype=torch.float32, device=dev)
    g = L.MsdaGrads()
    d_out = d_out.contiguous()
    d_query = torch.empty((B, Lq, C_), dtype=adt, device=dev)
    d_value = torch.empty((B, Lv, C_), dtype=adt, device=dev)
    d_ref = torch.empty((a.ref_batches, Lq, nL, 2), **f32) if want_ref_grad else None
    dw_q, db_q = torch.zeros((C_, 3 * tp), **f32), torch.zeros((3 * tp,), **f32)
    dw_v, db_v = torch.zeros((C_, C_), **f32), torch.zeros((C_,), **f32)
    dw_o, db_o = torch.zeros((C_, C_), **f32), torch.zeros((C_,), **f32)
    ws = torch.empty((int(lib.emrt_msda_fused_bwd_workspace_bytes(B, Lq, Lv, C_, M, nL, P, a.dtype)),), dtype=torch.uint8, device=dev)
    g.d_out, g.d_query, g.d_value, g.d_ref = _ptr(d_out), _ptr(d_query), _ptr(d_value), _ptr(d_ref)
    g.dw_query, g.db_query, g.dw_value, g.db_value, g.dw_out, g.db_out = (_ptr(t) for t in (dw_q, db_q, dw_v, db_v, dw_o, db_o))
    g.wq_cat, g.w_value_cast, g.w_out_cast, g.workspace = _ptr(wq_cat), _ptr(w_value_cast), _ptr(w_out_cast), _ptr(ws)
    L.check(lib.emrt_msda_fused_bwd(C.byref(a), C.byref(g), _stream()))
    return d_query, d_value, d_ref, dw_q, db_q, dw_v, db_v, dw_o, db_o
