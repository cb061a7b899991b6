"""torch.autograd.Function wrappers of the training path (SURVEY.md §8e cfg 4): the reference trains the whole
``EncoderDecoder`` through Paddle autograd (train.py:146-159); here every forward AND backward step of the glue around
``MSDeformableAttention`` is a kernel of libemrt_b200.so — torch's autograd engine only orders the calls and adds parameter
gradients into their (all-reduce bucket) ``.grad`` tensors.  Activations fp32 (parity path) or bf16; parameters and their
gradients fp32.  No arithmetic on activations is done by torch in this file: `.contiguous()`, views, `torch.cat` and dtype
casts of parameters are data movement."""
from __future__ import annotations

import torch

from . import _lib as L
from . import ops

_F32 = torch.float32


def _impl(x, impl):
    return impl if x.dtype == torch.bfloat16 else L.IMPL_SIMT


class LinearFn(torch.autograd.Function):
    """y = act(x @ W + b), W [in, out] (Paddle nn.Linear layout, fp32 master), act = ReLU or identity."""

    @staticmethod
    def forward(ctx, x, w, b, relu, impl):
        x = x.contiguous()
        fast = x.dtype == torch.bfloat16
        epi = L.EPI_RELU if relu else L.EPI_NONE
        bias = None if b is None else b.detach().float().contiguous()
        if fast and impl != L.IMPL_SIMT:
            wt = torch.empty((w.shape[1], w.shape[0]), dtype=torch.bfloat16, device=w.device)
            ops.pack_weight(w.detach().float().contiguous(), wt)
            y = ops.linear(x, wt, bias, w_transposed=True, epilogue=epi, impl=impl)
        else:
            y = ops.linear(x, w.detach().to(x.dtype).contiguous(), bias, epilogue=epi, impl=L.IMPL_SIMT)
        ctx.relu, ctx.impl, ctx.has_bias = relu, impl, b is not None
        ctx.save_for_backward(x, w, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        dy = dy.contiguous()
        if ctx.relu:
            dy = ops.relu_bwd(dy, y)
        impl = _impl(x, ctx.impl)
        dx = None
        if ctx.needs_input_grad[0]:
            # Paddle's [in, out] layout IS the K-major operand of the data-gradient GEMM (dx = dy W^T)
            dx = ops.linear(dy, w.detach().to(x.dtype).contiguous(), None, w_transposed=True, impl=impl).view(x.shape)
        K, N = w.shape
        dw = torch.zeros((K, N), dtype=_F32, device=x.device)
        db = torch.zeros((N,), dtype=_F32, device=x.device) if ctx.has_bias else None
        ops.linear_bwd_weight(x, dy, dw, db)
        return dx, dw, db, None, None


def linear(x, w, b, relu=False, impl=L.IMPL_AUTO):
    return LinearFn.apply(x, w, b, relu, impl)


class LayerNormFn(torch.autograd.Function):
    """y = LayerNorm(a + b) * gamma + beta (t_e_d.py:199-200,159-160,285-295)."""

    @staticmethod
    def forward(ctx, a, b, gamma, beta, eps):
        a, b = a.contiguous(), b.contiguous()
        g, bt = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        y = ops.residual_layernorm(b, a, g, bt, eps=eps)
        ctx.eps = eps
        ctx.save_for_backward(a, b, g)
        return y

    @staticmethod
    def backward(ctx, dy):
        a, b, g = ctx.saved_tensors
        N = a.shape[-1]
        dg = torch.zeros((N,), dtype=_F32, device=a.device)
        db = torch.zeros((N,), dtype=_F32, device=a.device)
        dz = ops.layernorm_bwd(a, b, g, dy.contiguous(), dg, db, eps=ctx.eps)
        return dz, dz, dg, db, None


def add_layernorm(a, b, norm, eps=1e-5):
    return LayerNormFn.apply(a, b, norm.weight, norm.bias, eps)


class AddFn(torch.autograd.Function):
    """a + b, b broadcast over the leading dimension when smaller (with_pos_embed, t_e_d.py:154-155; the layer's final add,
    :203).  The gradient of a broadcast addend is the batch sum of the incoming gradient."""

    @staticmethod
    def forward(ctx, a, b):
        a = a.contiguous()
        bb = b.to(a.dtype).contiguous()
        ctx.bshape, ctx.bdtype, ctx.bcast = b.shape, b.dtype, b.numel() != a.numel()
        return ops.add_bcast(a, bb)

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        db = None
        if ctx.needs_input_grad[1]:
            if ctx.bcast:
                n = 1
                for s in ctx.bshape:
                    n *= s
                db = ops.batch_sum(dy.view(dy.numel() // n, n)).view(ctx.bshape).to(ctx.bdtype)
            else:
                db = dy.view(ctx.bshape).to(ctx.bdtype)
        return (dy if ctx.needs_input_grad[0] else None), db


def add(a, b):
    return AddFn.apply(a, b)


class ConvBranchFn(torch.autograd.Function):
    """GELU(GroupNorm_l(conv3x3_l(src))) + src per level on the token layout (t_e_d.py:125-144,163-196)."""

    @staticmethod
    def forward(ctx, src, shapes, impl, eps, w0, w1, w2, g0, g1, g2, b0, b1, b2):
        src = src.contiguous()
        ws = [w.detach() for w in (w0, w1, w2)]
        gamma = torch.stack([g.detach().float() for g in (g0, g1, g2)]).contiguous()
        beta = torch.stack([b.detach().float() for b in (b0, b1, b2)]).contiguous()
        cimpl = L.IMPL_AUTO if (src.dtype == torch.bfloat16 and impl != L.IMPL_SIMT) else L.IMPL_SIMT
        conv = ops.conv3x3_tokens(src, ops.pack_conv3x3_weights(ws, src.dtype), shapes, impl=cimpl)
        out, stats = ops.groupnorm_gelu_residual(conv, src, gamma, beta, shapes, groups=32, eps=eps, return_stats=True)
        ctx.shapes, ctx.cimpl, ctx.eps = shapes, cimpl, eps
        ctx.save_for_backward(src, conv, stats, gamma, beta, w0, w1, w2)
        return out

    @staticmethod
    def backward(ctx, d_out):
        src, conv, stats, gamma, beta, w0, w1, w2 = ctx.saved_tensors
        d_out = d_out.contiguous()
        shapes = ctx.shapes
        dgamma, dbeta = torch.zeros_like(gamma), torch.zeros_like(beta)
        dconv = ops.groupnorm_bwd(conv, d_out, stats, gamma, beta, dgamma, dbeta, shapes, groups=32, eps=ctx.eps, gelu=True)
        # data gradient = the same conv on dconv with the taps flipped and (Cout, Cin) swapped
        flipped = [w.detach().flip(2, 3).transpose(0, 1).contiguous() for w in (w0, w1, w2)]
        dsrc = ops.conv3x3_tokens(dconv, ops.pack_conv3x3_weights(flipped, src.dtype), shapes, impl=ctx.cimpl)
        dsrc = ops.add_bcast(dsrc, d_out, out=dsrc)                                     # + the skip connection
        dw = ops.conv3x3_tokens_bwd_weight(src, dconv, shapes, impl=L.IMPL_AUTO if ctx.cimpl != L.IMPL_SIMT else L.IMPL_SIMT)
        return (dsrc, None, None, None, dw[0], dw[1], dw[2], dgamma[0], dgamma[1], dgamma[2], dbeta[0], dbeta[1], dbeta[2])


class GroupNormTokensFn(torch.autograd.Function):
    """GroupNorm(32, C) of one level's tokens [B, P, C] (input_proj.{l}.1, t_e_d.py:417-419)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        x = x.contiguous()
        g, b = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        out = torch.empty_like(x)
        _, stats = ops.groupnorm_tokens_into(x, g, b, out, 0, groups=32, eps=eps, return_stats=True)
        ctx.eps = eps
        ctx.save_for_backward(x, stats, g, b)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, stats, g, b = ctx.saved_tensors
        dg, db = torch.zeros_like(g), torch.zeros_like(b)
        dx = ops.groupnorm_bwd(x, dy.contiguous(), stats, g.view(1, -1), b.view(1, -1), dg.view(1, -1), db.view(1, -1),
                               [(x.shape[1], 1)], groups=32, eps=ctx.eps, gelu=False)
        return dx, dg, db, None


class TokensFn(torch.autograd.Function):
    """[B, C, *spatial] -> tokens [B, prod(spatial), C] (the flatten(2).transpose of t_e_d.py:446,469)."""

    @staticmethod
    def forward(ctx, x):
        ctx.spatial = tuple(x.shape[2:])
        return ops.nchw_to_tokens(x)

    @staticmethod
    def backward(ctx, dy):
        return ops.tokens_to_nchw(dy.contiguous(), ctx.spatial)


class SelfAttentionCoreFn(torch.autograd.Function):
    """softmax(scale q k^T) v per head on the fused [q | k] projection and v (layers.py:282-301)."""

    @staticmethod
    def forward(ctx, qk, v, num_heads, scale):
        qk, v = qk.contiguous(), v.contiguous()
        C_ = v.shape[-1]
        ctx.num_heads, ctx.scale = num_heads, scale
        ctx.save_for_backward(qk, v)
        return ops.mha_small(qk[..., :C_], qk[..., C_:], v, num_heads, scale)

    @staticmethod
    def backward(ctx, d_out):
        qk, v = ctx.saved_tensors
        C_ = v.shape[-1]
        dqk, dv = ops.mha_small_bwd(qk[..., :C_], qk[..., C_:], v, d_out, ctx.num_heads, ctx.scale)
        return dqk, dv, None, None


class PosEmbedFn(torch.autograd.Function):
    """pos[token] = sine[token] + level_embed[level(token)] (t_e_d.py:444-447): `sine` is the cached host constant."""

    @staticmethod
    def forward(ctx, level_embed, sine, shapes, dtype):
        le = level_embed.detach().float().contiguous()
        parts, off = [], 0
        for l, (h, w) in enumerate(shapes):
            parts.append(ops.add_bcast(sine[off:off + h * w].contiguous(), le[l].contiguous()))
            off += h * w
        ctx.shapes = shapes
        return torch.cat(parts, 0)[None].to(dtype)

    @staticmethod
    def backward(ctx, d_pos):
        d = d_pos[0].contiguous()
        rows, off = [], 0
        for (h, w) in ctx.shapes:
            rows.append(ops.column_sum(d[off:off + h * w].contiguous()))
            off += h * w
        return torch.stack(rows), None, None, None


class ReferencePointsFn(torch.autograd.Function):
    """sigmoid(Linear(query_pos_embed)) replicated over the levels (t_e_d.py:464-467, valid_ratios == 1) -> [1, Nq, L, 2]."""

    @staticmethod
    def forward(ctx, qpe, w, b, n_levels):
        x = qpe.detach().float().contiguous()
        lin = ops.linear(x, w.detach().float().contiguous(), b.detach().float().contiguous(), impl=L.IMPL_SIMT)
        y = ops.sigmoid(lin)
        ctx.n_levels = n_levels
        ctx.save_for_backward(x, w, y)
        return y[None, :, None, :].expand(1, y.shape[0], n_levels, 2).contiguous()

    @staticmethod
    def backward(ctx, d_ref):
        x, w, y = ctx.saved_tensors
        dy = ops.batch_sum(d_ref[0].float().permute(1, 0, 2).contiguous())                   # sum over the levels -> [Nq, 2]
        dlin = ops.sigmoid_bwd(dy.contiguous(), y)
        dx = ops.linear(dlin, w.detach().float().contiguous(), None, w_transposed=True, impl=L.IMPL_SIMT)
        dw = torch.zeros(tuple(w.shape), dtype=_F32, device=x.device)
        db = torch.zeros((w.shape[1],), dtype=_F32, device=x.device)
        ops.linear_bwd_weight(x, dlin, dw, db)
        return dx, dw, db, None
