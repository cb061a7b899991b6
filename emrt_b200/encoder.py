"""Host-side mirror of the reference's encoder around the MSDA module:
  * ``TransformerEncoderLayer`` — src/models/EMRT_utils/transformer_encoder_decoder.py:109-204
  * ``TransformerEncoder``      — :207-239
Same constructor arguments, sub-layer names (state-dict keys: ``self_attn.*``, ``norm1``, ``linear1``, ``linear2``,
``norm2``, ``conv{0,1,2}.0.weight`` [Cout,Cin,3,3], ``conv{l}.1.weight/bias`` GroupNorm) and forward signatures.
Dropout is the identity (eval-mode semantics).  Inference runs the fused kernels; when gradients are required the same
layer runs as a differentiable composition whose forward and backward steps are all kernels (emrt_b200/autograd.py).  All
arithmetic runs in libemrt_b200.so on the token layout [B, Lv, C], so the reference's seq2_2D / flatten / transpose /
concat copies (:163-196) do not exist here.
"""
from __future__ import annotations

import math
import os

import torch
from torch import nn

from . import _lib as L
from . import ops
from .msda import MSDeformableAttention, PaddleLinear, shapes_to_host
from .refpoints import get_reference_points


class _Norm(nn.Module):
    def __init__(self, n):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(n))
        self.bias = nn.Parameter(torch.zeros(n))


class _ConvGN(nn.Module):
    """nn.Sequential(Conv2D(C, C, 3, padding=1, bias_attr=False), GroupNorm(32, C), GELU()) parameter container with the
    reference's key names ``0.weight``, ``1.weight``, ``1.bias``."""

    def __init__(self, c):
        super().__init__()
        conv = nn.Module()
        conv.weight = nn.Parameter(torch.empty(c, c, 3, 3))
        nn.init.kaiming_uniform_(conv.weight, a=math.sqrt(5))
        self.add_module("0", conv)
        self.add_module("1", _Norm(c))


class TransformerEncoderLayer(nn.Module):
    def __init__(self, d_model=256, n_head=8, dim_feedforward=1024, dropout=0.1, activation="relu", n_levels=4,
                 n_points=4, weight_attr=None, bias_attr=None):
        super().__init__()
        if activation != "relu":
            raise L.EmrtError("emrt_b200.TransformerEncoderLayer implements the reference's activation='relu' only")
        self.d_model, self.n_levels = d_model, n_levels
        self.self_attn = MSDeformableAttention(d_model, n_head, n_levels, n_points)
        self.norm1 = _Norm(d_model)
        self.linear1 = PaddleLinear(d_model, dim_feedforward)
        self.linear2 = PaddleLinear(dim_feedforward, d_model)
        self.norm2 = _Norm(d_model)
        for l in range(3):                                   # the reference hard-codes conv0..conv2 (:125-144)
            self.add_module(f"conv{l}", _ConvGN(d_model))
        nn.init.xavier_uniform_(self.linear1.weight)
        nn.init.xavier_uniform_(self.linear2.weight)
        self.gemm_impl = L.IMPL_AUTO
        self.fuse_ffn2_norm2 = True          # tests may switch the fused linear2 + norm2 + conv-branch epilogue off
        self.fuse_ffn = True                 # ... and the whole-FFN kernel (falls back to linear1 -> fused linear2 epilogue)
        self.conv_stats_fused = True         # GroupNorm statistics from the conv's epilogue (else the separate statistics kernel)
        # SMs the conv takes on a side stream while the gather runs (0: one stream, in order); EMRT_OVERLAP_CONV overrides
        self.overlap_conv = int(os.environ.get("EMRT_OVERLAP_CONV", "64"))
        self._packed = None

    def _version(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _packed_weights(self, dtype):
        ver = (self._version(), dtype)
        if self._packed is not None and self._packed[0] == ver:
            return self._packed[1]
        convs = [getattr(self, f"conv{l}") for l in range(3)]
        f32 = lambda t: t.detach().float().contiguous()
        pk = dict(conv_w=ops.pack_conv3x3_weights([getattr(c, "0").weight for c in convs], dtype),
                  gn_w=torch.stack([f32(getattr(c, "1").weight) for c in convs]).contiguous(),
                  gn_b=torch.stack([f32(getattr(c, "1").bias) for c in convs]).contiguous(),
                  n1w=f32(self.norm1.weight), n1b=f32(self.norm1.bias), n2w=f32(self.norm2.weight),
                  n2b=f32(self.norm2.bias), b1=f32(self.linear1.bias), b2=f32(self.linear2.bias))
        if dtype == torch.bfloat16:
            C_, F_ = self.d_model, self.linear1.weight.shape[1]
            w1 = torch.empty((F_, C_), dtype=torch.bfloat16, device=self.linear1.weight.device)
            w2 = torch.empty((C_, F_), dtype=torch.bfloat16, device=self.linear1.weight.device)
            ops.pack_weight(f32(self.linear1.weight), w1)
            ops.pack_weight(f32(self.linear2.weight), w2)
            pk.update(w1=w1, w2=w2)
        self._packed = (ver, pk)
        return pk

    _emrt_train = False

    def train(self, mode: bool = True):
        """`.train()` (train.py:139) switches to the differentiable path, `.eval()` back to the fused inference kernels.  A
        freshly constructed layer runs the inference path (nn.Module's own default `training = True` does not count)."""
        self._emrt_train = bool(mode)
        return super().train(mode)

    def _wants_grad(self, *tensors):
        return self._emrt_train and torch.is_grad_enabled() and (any(t is not None and t.requires_grad for t in tensors)
                                                                 or any(p.requires_grad for p in self.parameters()))

    def forward(self, src, reference_points, spatial_shapes, src_mask=None, pos_embed=None):
        if self._wants_grad(src, pos_embed):
            return self._forward_train(src, reference_points, spatial_shapes, src_mask, pos_embed)
        with torch.no_grad():
            return self._forward_eval(src, reference_points, spatial_shapes, src_mask, pos_embed)

    def _forward_train(self, src, reference_points, spatial_shapes, src_mask=None, pos_embed=None):
        """The same layer (t_e_d.py:184-204) as a differentiable composition: every step is an autograd.Function whose
        forward and backward are kernels of libemrt_b200.so (emrt_b200/autograd.py).  Dropout is the identity
        (the mirrors implement eval-mode dropout; train with dropout = 0 for parity with the reference's step)."""
        from . import autograd as A
        shapes = shapes_to_host(spatial_shapes)
        if len(shapes) != 3:
            raise L.EmrtError("the reference's encoder layer is written for 3 feature levels (conv0..conv2)")
        impl = self.gemm_impl if src.dtype == torch.bfloat16 else L.IMPL_SIMT
        convs = [getattr(self, f"conv{l}") for l in range(3)]
        branch = A.ConvBranchFn.apply(src, shapes, impl, 1e-5, *[getattr(c, "0").weight for c in convs],
                                      *[getattr(c, "1").weight for c in convs], *[getattr(c, "1").bias for c in convs])
        q = src if pos_embed is None else A.add(src, pos_embed)                                     # with_pos_embed (:198)
        src2 = self.self_attn(q, reference_points, src, shapes, src_mask)
        x = A.add_layernorm(src, src2, self.norm1)                                                  # :199-200
        h = A.linear(x, self.linear1.weight, self.linear1.bias, relu=True, impl=impl)               # :157-158
        f = A.linear(h, self.linear2.weight, self.linear2.bias, impl=impl)
        y = A.add_layernorm(x, f, self.norm2)                                                       # :159-160
        return A.add(y, branch)                                                                     # :203

    def _forward_eval(self, src, reference_points, spatial_shapes, src_mask=None, pos_embed=None):
        shapes = shapes_to_host(spatial_shapes)
        if len(shapes) != 3:
            raise L.EmrtError("the reference's encoder layer is written for 3 feature levels (conv0..conv2)")
        src = src.contiguous()
        pk = self._packed_weights(src.dtype)
        fast = src.dtype == torch.bfloat16
        impl = self.gemm_impl if fast else L.IMPL_SIMT
        # conv branch (:185-196): conv3x3 -> GroupNorm(32) -> GELU, + skip, on the token layout
        # (GroupNorm + GELU + skip are applied inside the last LayerNorm pass below: the branch tensor never exists)
        conv = gn_stats = None
        tc_conv = fast and impl != L.IMPL_SIMT and self.d_model == 256 and self.conv_stats_fused
        if tc_conv and self.overlap_conv > 0 and not torch.cuda.is_current_stream_capturing():
            return self._forward_eval_overlapped(src, reference_points, shapes, src_mask, pos_embed, pk)
        if tc_conv:
            try:        # the conv's epilogue leaves the GroupNorm statistics: no separate pass over its output
                conv, gn_stats = ops.conv3x3_tokens_stats(src, pk["conv_w"], shapes, groups=32)
            except L.EmrtError as exc:
                if getattr(exc, "status", 0) != L.ERR_UNSUPPORTED:      # a shape the tcgen05 conv does not tile: two-kernel form
                    raise
        if conv is None:
            conv = ops.conv3x3_tokens(src, pk["conv_w"], shapes, impl=L.IMPL_AUTO if impl != L.IMPL_SIMT else L.IMPL_SIMT)
            gn_stats = ops.groupnorm_stats(conv, shapes, groups=32)
        # self attention (:198) + norm1 (:199-200)
        # (with_pos_embed is folded into the query projection, norm1 into the output projection: msda.py)
        x = self.self_attn(src, reference_points, src, shapes, src_mask, query_pos=pos_embed,
                           residual_norm=(src, pk["n1w"], pk["n1b"]))
        # ffn (:157-160) + the layer's final add of the conv branch (:203)
        if fast and impl != L.IMPL_SIMT and self.fuse_ffn and self.d_model == 256 and pk["w1"].shape[0] % 128 == 0:
            # linear1 + ReLU + linear2 + residual + norm2 + the conv branch + the layer's final add in ONE kernel: the hidden
            # activations never leave the SM (emrt_ffn_fused_fwd; :157-160,187-189,203)
            return ops.ffn_fused(x, pk["w1"], pk["b1"], pk["w2"], pk["b2"], pk["n2w"], pk["n2b"],
                                 gn_branch=dict(conv=conv, skip=src, stats=gn_stats, gamma=pk["gn_w"], beta=pk["gn_b"],
                                                shapes=shapes, groups=32, eps=1e-5))
        if fast:
            h = ops.linear(x, pk["w1"], pk["b1"], w_transposed=True, epilogue=L.EPI_RELU, impl=impl)
            if impl != L.IMPL_SIMT and self.d_model == 256 and self.fuse_ffn2_norm2:
                # linear2 + residual + norm2 + the conv branch + the layer's final add in ONE kernel: the LayerNorm runs on the
                # fp32 accumulator and GELU(GroupNorm(conv)) + src is added in the same epilogue (:157-160,187-189,203)
                return ops.linear(h, pk["w2"], pk["b2"], w_transposed=True, impl=impl, epilogue=L.EPI_RESIDUAL_LN, residual=x,
                                  ln_gamma=pk["n2w"], ln_beta=pk["n2b"],
                                  gn_branch=dict(conv=conv, skip=src, stats=gn_stats, gamma=pk["gn_w"], beta=pk["gn_b"],
                                                 shapes=shapes, groups=32, eps=1e-5))
            f = ops.linear(h, pk["w2"], pk["b2"], w_transposed=True, impl=impl)
        else:
            h = ops.linear(x, self.linear1.weight.detach(), pk["b1"], epilogue=L.EPI_RELU, impl=L.IMPL_SIMT)
            f = ops.linear(h, self.linear2.weight.detach(), pk["b2"], impl=L.IMPL_SIMT)
        return ops.residual_layernorm_gn(f, x, pk["n2w"], pk["n2b"], conv, src, gn_stats, pk["gn_w"], pk["gn_b"], shapes,
                                         groups=32, out=f)


    def _forward_eval_overlapped(self, src, reference_points, shapes, src_mask, pos_embed, pk):
        """The bf16 layer with its two independent branches (conv :185-196, attention :198-200) on two streams for the one
        stretch where that pays: the 3x3 convolution — tensor-pipe bound, and power-limited when it owns all 148 SMs (it
        costs 28 % less SM time on a share of them) — starts on `overlap_conv` SMs of a high-priority side stream at the
        moment the sampling gather — bound by the shared-memory pipe — is launched on the main stream; when the conv is done
        its SMs go to the gather's remaining CTAs.  conv + gather: 951 -> 909 us per layer
        (profiles/r3t_conv_gather_overlap.txt).  Same kernels, same results bit for bit; the main stream waits for the
        conv before the FFN kernel that consumes it."""
        main = torch.cuda.current_stream()
        side = _side_stream(src.device)
        conv, gn_stats = ops.conv3x3_stats_buffers(src, shapes)          # allocated on the stream that consumes them
        start = torch.cuda.Event()
        x = self.self_attn(src, reference_points, src, shapes, src_mask, query_pos=pos_embed,
                           residual_norm=(src, pk["n1w"], pk["n1b"]), gather_start_event=start)
        try:
            with torch.cuda.stream(side):
                side.wait_event(start)
                ops.conv3x3_tokens_stats(src, pk["conv_w"], shapes, groups=32, out=conv, stats=gn_stats, max_ctas=self.overlap_conv)
                done = torch.cuda.Event()
                done.record(side)
            main.wait_event(done)
        except L.EmrtError as exc:
            if getattr(exc, "status", 0) != L.ERR_UNSUPPORTED:          # a shape the tcgen05 conv does not tile: two-kernel form
                raise
            conv = ops.conv3x3_tokens(src, pk["conv_w"], shapes, impl=L.IMPL_AUTO)
            gn_stats = ops.groupnorm_stats(conv, shapes, groups=32)
        gn = dict(conv=conv, skip=src, stats=gn_stats, gamma=pk["gn_w"], beta=pk["gn_b"], shapes=shapes, groups=32, eps=1e-5)
        if self.fuse_ffn and pk["w1"].shape[0] % 128 == 0:
            return ops.ffn_fused(x, pk["w1"], pk["b1"], pk["w2"], pk["b2"], pk["n2w"], pk["n2b"], gn_branch=gn)
        h = ops.linear(x, pk["w1"], pk["b1"], w_transposed=True, epilogue=L.EPI_RELU, impl=self.gemm_impl)
        if self.fuse_ffn2_norm2:
            return ops.linear(h, pk["w2"], pk["b2"], w_transposed=True, impl=self.gemm_impl, epilogue=L.EPI_RESIDUAL_LN, residual=x,
                              ln_gamma=pk["n2w"], ln_beta=pk["n2b"], gn_branch=gn)
        f = ops.linear(h, pk["w2"], pk["b2"], w_transposed=True, impl=self.gemm_impl)
        return ops.residual_layernorm_gn(f, x, pk["n2w"], pk["n2b"], conv, src, gn_stats, pk["gn_w"], pk["gn_b"], shapes,
                                         groups=32, out=f)


_SIDE_STREAMS = {}


def _side_stream(device):
    """One high-priority side stream per device: when the conv and the gather become runnable together, the block scheduler
    places the conv's few CTAs first (each needs a whole SM's shared memory) and the gather fills the rest."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=key, priority=-1)
    return _SIDE_STREAMS[key]


class TransformerEncoder(nn.Module):
    """transformer_encoder_decoder.py:207-239: reference points once (cached constant), then the layers."""

    def __init__(self, encoder_layer, num_layers):
        super().__init__()
        import copy
        self.layers = nn.ModuleList([copy.deepcopy(encoder_layer) for _ in range(num_layers)])   # _get_clones
        self.num_layers = num_layers

    get_reference_points = staticmethod(get_reference_points)

    def forward(self, src, spatial_shapes, src_mask=None, pos_embed=None, valid_ratios=None):
        output = src
        reference_points = get_reference_points(spatial_shapes, valid_ratios, device=src.device)
        for layer in self.layers:
            output = layer(output, reference_points, spatial_shapes, src_mask, pos_embed)
        return output
