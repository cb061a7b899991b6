"""Host-side mirror of the reference's decoder and of EncoderDecoder:
  * ``MultiHeadAttention``       — src/models/EMRT_utils/layers.py:144-311 (fused in_proj_weight [C, 3C], out_proj)
  * ``TransformerDecoderLayer``  — src/models/EMRT_utils/transformer_encoder_decoder.py:242-295
  * ``TransformerDecoder``       — :298-334
  * ``EncoderDecoder``           — :337-473 (input_proj 1x1 conv + GroupNorm, sine position embedding + level embed,
                                    query embeddings, reference-point Linear + sigmoid, encoder, decoder)
Same constructor arguments, parameter names (state-dict keys) and forward signatures.  Inference runs the fused kernels;
when gradients are required (cfg 4: the reference trains the whole EncoderDecoder) every layer runs as a differentiable
composition whose forward AND backward steps are kernels (emrt_b200/autograd.py).  All arithmetic on
activations runs in libemrt_b200.so; what is computed on the host is input-independent and cached: the sine position
embedding (position_encoding.py:51-75 with an all-ones mask) + level embedding, and the decoder reference points
sigmoid(Linear(query_pos_embed.weight)) (:466) — functions of the weights and the level shapes only.
"""
from __future__ import annotations

import copy
import math
from typing import Sequence

import numpy as np
import torch
from torch import nn

from . import _lib as L
from . import ops
from .encoder import TransformerEncoder, TransformerEncoderLayer, _Norm
from .msda import MSDeformableAttention, PaddleLinear, shapes_to_host


def position_embedding_sine_host(h: int, w: int, num_pos_feats=128, temperature=10000.0, offset=-0.5, eps=1e-6,
                                 scale=2.0 * math.pi) -> np.ndarray:
    """PositionEmbedding.forward (position_encoding.py:51-75) for an all-ones mask, float32 -> [h*w, 2*num_pos_feats]
    (pos_y | pos_x), token-major (the `.flatten(2).transpose([0, 2, 1])` of t_e_d.py:446 already applied)."""
    f = np.float32
    y_embed = np.cumsum(np.ones((h, w), f), 0, dtype=f)
    x_embed = np.cumsum(np.ones((h, w), f), 1, dtype=f)
    y_embed = (y_embed + f(offset)) / (y_embed[-1:, :] + f(eps)) * f(scale)
    x_embed = (x_embed + f(offset)) / (x_embed[:, -1:] + f(eps)) * f(scale)
    dim_t = (2 * (np.arange(num_pos_feats) // 2)).astype(f)
    dim_t = np.power(f(temperature), dim_t / f(num_pos_feats)).astype(f)
    pos_x = x_embed[..., None] / dim_t
    pos_y = y_embed[..., None] / dim_t
    pos_x = np.stack((np.sin(pos_x[..., 0::2]), np.cos(pos_x[..., 1::2])), axis=3).reshape(h, w, -1)
    pos_y = np.stack((np.sin(pos_y[..., 0::2]), np.cos(pos_y[..., 1::2])), axis=3).reshape(h, w, -1)
    return np.concatenate((pos_y, pos_x), axis=2).reshape(h * w, -1).astype(f)


TRAINABLE = True      # the whole EncoderDecoder has a backward (emrt_b200/train.py::build_train_step uses it for cfg 4)


class MultiHeadAttention(nn.Module):
    """Parameter container + forward of layers.py:144-311 (self-attention use: q = k = tgt + pos, value = tgt)."""

    def __init__(self, embed_dim, num_heads, dropout=0.0):
        super().__init__()
        self.embed_dim, self.num_heads = embed_dim, num_heads
        self.head_dim = embed_dim // num_heads
        assert self.head_dim * num_heads == self.embed_dim, "embed_dim must be divisible by num_heads"
        self.in_proj_weight = nn.Parameter(torch.empty(embed_dim, 3 * embed_dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dim))
        self.out_proj = PaddleLinear(embed_dim, embed_dim)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.xavier_uniform_(self.out_proj.weight)


class TransformerDecoderLayer(nn.Module):
    def __init__(self, d_model=256, n_head=8, dim_feedforward=1024, dropout=0.1, activation="relu", n_levels=3,
                 n_points=4, weight_attr=None, bias_attr=None):
        super().__init__()
        if activation != "relu":
            raise L.EmrtError("emrt_b200.TransformerDecoderLayer implements the reference's activation='relu' only")
        self.d_model, self.n_head = d_model, n_head
        self.self_attn = MultiHeadAttention(d_model, n_head, dropout=dropout)
        self.norm1 = _Norm(d_model)
        self.cross_attn = MSDeformableAttention(d_model, n_head, n_levels, n_points)
        self.norm2 = _Norm(d_model)
        self.linear1 = PaddleLinear(d_model, dim_feedforward)
        self.linear2 = PaddleLinear(dim_feedforward, d_model)
        self.norm3 = _Norm(d_model)
        nn.init.xavier_uniform_(self.linear1.weight)
        nn.init.xavier_uniform_(self.linear2.weight)
        self.gemm_impl = L.IMPL_AUTO
        self._packed = None

    def _packed_weights(self, dtype):
        ver = (tuple((p.data_ptr(), p._version) for p in self.parameters()), dtype)
        if self._packed is not None and self._packed[0] == ver:
            return self._packed[1]
        C_ = self.d_model
        f32 = lambda t: t.detach().float().contiguous()
        sa = self.self_attn
        pk = dict(b_qk=f32(sa.in_proj_bias[:2 * C_]), b_v=f32(sa.in_proj_bias[2 * C_:]), b_o=f32(sa.out_proj.bias),
                  b1=f32(self.linear1.bias), b2=f32(self.linear2.bias))
        for i, n in ((1, self.norm1), (2, self.norm2), (3, self.norm3)):
            pk[f"n{i}w"], pk[f"n{i}b"] = f32(n.weight), f32(n.bias)
        srcs = dict(w_qk=sa.in_proj_weight[:, :2 * C_], w_v=sa.in_proj_weight[:, 2 * C_:], w_o=sa.out_proj.weight,
                    w1=self.linear1.weight, w2=self.linear2.weight)
        for name, w in srcs.items():
            w = f32(w)
            if dtype == torch.bfloat16:       # K-major bf16 [out, in] operands for the tcgen05 GEMMs
                dst = torch.empty((w.shape[1], w.shape[0]), dtype=torch.bfloat16, device=w.device)
                pk[name] = ops.pack_weight(w, dst)
            else:
                pk[name] = w                  # Paddle [in, out] layout, fp32 SIMT path
        self._packed = (ver, pk)
        return pk

    _emrt_train = False

    def train(self, mode: bool = True):
        self._emrt_train = bool(mode)          # see TransformerEncoderLayer.train
        return super().train(mode)

    def forward(self, tgt, reference_points, memory, memory_spatial_shapes, memory_mask=None, query_pos_embed=None):
        if self._emrt_train and torch.is_grad_enabled() and (any(t is not None and t.requires_grad for t in (tgt, memory, reference_points, query_pos_embed))
                                        or any(p.requires_grad for p in self.parameters())):
            return self._forward_train(tgt, reference_points, memory, memory_spatial_shapes, memory_mask, query_pos_embed)
        with torch.no_grad():
            return self._forward_eval(tgt, reference_points, memory, memory_spatial_shapes, memory_mask, query_pos_embed)

    def _forward_train(self, tgt, reference_points, memory, memory_spatial_shapes, memory_mask=None, query_pos_embed=None):
        """t_e_d.py:282-295 as a differentiable composition of kernel-backed autograd Functions (emrt_b200/autograd.py)."""
        from . import autograd as A
        shapes = shapes_to_host(memory_spatial_shapes)
        impl = self.gemm_impl if tgt.dtype == torch.bfloat16 else L.IMPL_SIMT
        C_, M = self.d_model, self.n_head
        sa = self.self_attn
        with_pos = lambda t: t if query_pos_embed is None else A.add(t, query_pos_embed)
        # self attention (layers.py:221-234,282-301): q = k = tgt + pos, value = tgt; in_proj_weight [C, 3C] sliced per use
        qk = A.linear(with_pos(tgt), sa.in_proj_weight[:, :2 * C_], sa.in_proj_bias[:2 * C_], impl=impl)
        v = A.linear(tgt, sa.in_proj_weight[:, 2 * C_:], sa.in_proj_bias[2 * C_:], impl=impl)
        att = A.SelfAttentionCoreFn.apply(qk, v, M, float(C_ // M) ** -0.5)
        tgt2 = A.linear(att, sa.out_proj.weight, sa.out_proj.bias, impl=impl)
        tgt = A.add_layernorm(tgt, tgt2, self.norm1)
        tgt2 = self.cross_attn(with_pos(tgt), reference_points, memory, shapes, memory_mask)
        tgt = A.add_layernorm(tgt, tgt2, self.norm2)
        h = A.linear(tgt, self.linear1.weight, self.linear1.bias, relu=True, impl=impl)
        f = A.linear(h, self.linear2.weight, self.linear2.bias, impl=impl)
        return A.add_layernorm(tgt, f, self.norm3)

    def _forward_eval(self, tgt, reference_points, memory, memory_spatial_shapes, memory_mask=None, query_pos_embed=None):
        shapes = shapes_to_host(memory_spatial_shapes)
        tgt = tgt.contiguous()
        fast = tgt.dtype == torch.bfloat16
        impl = self.gemm_impl if fast else L.IMPL_SIMT
        pk = self._packed_weights(tgt.dtype)
        C_, M = self.d_model, self.n_head
        lin = lambda x, w, b, **kw: ops.linear(x, pk[w], pk[b], w_transposed=fast, impl=impl, **kw)
        pos = None if query_pos_embed is None else query_pos_embed.to(tgt.dtype).contiguous()
        Nq = tgt.shape[1]
        tc = fast and impl != L.IMPL_SIMT
        # with_pos_embed folded into the projections ((tgt + pos) W = tgt W + pos W in one GEMM) on the tcgen05 path
        fold = tc and pos is not None and pos.numel() == Nq * C_
        x2 = dict(x2=ops.cyclic_rows_cached(query_pos_embed), x2_period=Nq) if fold else {}
        with_pos = lambda t: t if (pos is None or fold) else ops.add_bcast(t, pos)

        def lin_norm(x, w, b, res, nw, nb):
            """LayerNorm(x W + b + res): one kernel on the tcgen05 path (EMRT_EPI_RESIDUAL_LN)."""
            if tc and C_ == 256:
                return ops.linear(x, pk[w], pk[b], w_transposed=True, impl=impl, epilogue=L.EPI_RESIDUAL_LN, residual=res,
                                  ln_gamma=pk[nw], ln_beta=pk[nb])
            y = lin(x, w, b)
            return ops.residual_layernorm(y, res, pk[nw], pk[nb], out=y)

        # self attention over the query tokens (layers.py:282-301): q = k = tgt + pos, value = tgt
        qk = lin(with_pos(tgt), "w_qk", "b_qk", **x2)
        v = lin(tgt, "w_v", "b_v")
        att = ops.mha_small(qk[..., :C_], qk[..., C_:], v, M, float(C_ // M) ** -0.5)
        tgt = lin_norm(att, "w_o", "b_o", tgt, "n1w", "n1b")
        # cross attention into the encoder memory (+ norm2)
        tgt = self.cross_attn(tgt, reference_points, memory, shapes, memory_mask, query_pos=pos,
                              residual_norm=(tgt, pk["n2w"], pk["n2b"]))
        # ffn (+ norm3)
        h = lin(tgt, "w1", "b1", epilogue=L.EPI_RELU)
        return lin_norm(h, "w2", "b2", tgt, "n3w", "n3b")


class TransformerDecoder(nn.Module):
    def __init__(self, decoder_layer, num_layers, return_intermediate=False):
        super().__init__()
        self.layers = nn.ModuleList([copy.deepcopy(decoder_layer) for _ in range(num_layers)])
        self.num_layers = num_layers
        self.return_intermediate = return_intermediate

    def forward(self, tgt, memory, reference_points, memory_spatial_shapes, memory_mask=None, query_pos_embed=None,
                valid_ratios=None):
        output = tgt
        intermediate = []
        for layer in self.layers:
            output = layer(output, reference_points, memory, memory_spatial_shapes, memory_mask, query_pos_embed)
            if self.return_intermediate:
                intermediate.append(output)
        if self.return_intermediate:
            return torch.stack(intermediate)
        return output.unsqueeze(0)


class _InputProj(nn.Module):
    """nn.Sequential(Conv2D(Cin, C, 1), GroupNorm(32, C)) parameter container (keys ``0.weight``, ``0.bias``,
    ``1.weight``, ``1.bias``)."""

    def __init__(self, cin, c):
        super().__init__()
        conv = nn.Module()
        conv.weight = nn.Parameter(torch.empty(c, cin, 1, 1))
        conv.bias = nn.Parameter(torch.zeros(c))
        nn.init.xavier_uniform_(conv.weight)
        self.add_module("0", conv)
        self.add_module("1", _Norm(c))


class _Embedding(nn.Module):
    def __init__(self, n, c):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(n, c))


class EncoderDecoder(nn.Module):
    def __init__(self, num_queries=110, position_embed_type="sine", return_intermediate_dec=False,
                 backbone_num_channels=(512, 1024, 2048), num_feature_levels=3, nclass=6, num_encoder_points=4,
                 num_decoder_points=4, hidden_dim=256, nhead=8, num_encoder_layers=6, num_decoder_layers=4,
                 dim_feedforward=1024, dropout=0.1, activation="relu", lr_mult=0.1, weight_attr=None, bias_attr=None):
        super().__init__()
        if position_embed_type != "sine":
            raise L.EmrtError("emrt_b200.EncoderDecoder implements position_embed_type='sine' (what EMRT uses)")
        if len(backbone_num_channels) != num_feature_levels:
            raise L.EmrtError("extra stride-2 input_proj levels (t_e_d.py:380-387) are not used by EMRT and not built")
        self.hidden_dim, self.nhead, self.num_feature_levels = hidden_dim, nhead, num_feature_levels
        self.input_proj_channel_major = True    # input_proj reads the NCHW maps directly (tests may force the transposed copy)
        self.encoder = TransformerEncoder(TransformerEncoderLayer(hidden_dim, nhead, dim_feedforward, dropout, activation,
                                                                  num_feature_levels, num_encoder_points), num_encoder_layers)
        self.decoder = TransformerDecoder(TransformerDecoderLayer(hidden_dim, nhead, dim_feedforward, dropout, activation,
                                                                  num_feature_levels, num_decoder_points),
                                          num_decoder_layers, return_intermediate_dec)
        self.level_embed = _Embedding(num_feature_levels, hidden_dim)
        self.tgt_embed = _Embedding(num_queries, hidden_dim)            # present in the checkpoint, never used (:368)
        self.query_pos_embed = _Embedding(num_queries, hidden_dim)
        self.reference_points = PaddleLinear(hidden_dim, 2)
        self.input_proj = nn.ModuleList([_InputProj(c, hidden_dim) for c in backbone_num_channels])
        self._const = None

    def _constants(self, shapes, device, dtype):
        """Input-independent tensors of the forward: position + level embedding [1, Lv, C], decoder reference points
        [1, Nq, L, 2], query position embedding [1, Nq, C], packed input_proj weights."""
        ver = (shapes, str(device), dtype, tuple((p.data_ptr(), p._version) for p in self.parameters()))
        if self._const is not None and self._const[0] == ver:
            return self._const[1]
        C_ = self.hidden_dim
        lvl = self.level_embed.weight.detach().float().cpu().numpy()
        pos = np.concatenate([position_embedding_sine_host(h, w, C_ // 2) + lvl[l][None] for l, (h, w) in enumerate(shapes)], 0)
        qpe = self.query_pos_embed.weight.detach().float().cpu().numpy()
        rp = qpe @ self.reference_points.weight.detach().float().cpu().numpy() + self.reference_points.bias.detach().float().cpu().numpy()
        rp = (1.0 / (1.0 + np.exp(-rp.astype(np.float64)))).astype(np.float32)                        # F.sigmoid (:466)
        rp = np.ascontiguousarray(np.broadcast_to(rp[None, :, None, :], (1, rp.shape[0], len(shapes), 2)))
        c = dict(pos=torch.from_numpy(pos[None]).to(device).to(dtype).contiguous(),
                 ref_dec=torch.from_numpy(rp).to(device),
                 qpos=torch.from_numpy(qpe[None]).to(device).to(dtype).contiguous(), w=[], b=[], gw=[], gb=[])
        for ip in self.input_proj:
            conv, gn = getattr(ip, "0"), getattr(ip, "1")
            w_kn = conv.weight.detach().float().reshape(conv.weight.shape[0], -1).t().contiguous()     # [Cin, C] = [in, out]
            if dtype == torch.bfloat16:
                dst = torch.empty((w_kn.shape[1], w_kn.shape[0]), dtype=torch.bfloat16, device=w_kn.device)
                c["w"].append(ops.pack_weight(w_kn, dst))
            else:
                c["w"].append(w_kn)
            c["b"].append(conv.bias.detach().float().contiguous())
            c["gw"].append(gn.weight.detach().float().contiguous())
            c["gb"].append(gn.bias.detach().float().contiguous())
        self._const = (ver, c)
        return c

    @torch.no_grad()
    def project_inputs(self, src_feats: Sequence[torch.Tensor]):
        """input_proj (:417-419) + flatten / concat (:434-452): 1x1 conv as a GEMM on tokens, GroupNorm written straight
        into the level's token slot.  -> (src [B, Lv, C], level shapes, the cached constants of `_constants`)."""
        x0 = src_feats[0]
        B, dtype, dev = x0.shape[0], x0.dtype, x0.device
        shapes = tuple((int(f.shape[2]), int(f.shape[3])) for f in src_feats)
        Lv = sum(h * w for h, w in shapes)
        c = self._constants(shapes, dev, dtype)
        fast = dtype == torch.bfloat16
        src = torch.empty((B, Lv, self.hidden_dim), dtype=dtype, device=dev)
        off = 0
        for l, f in enumerate(src_feats):
            hw = shapes[l][0] * shapes[l][1]
            if fast and self.input_proj_channel_major and hw % 128 == 0 and f.shape[1] % 64 == 0 and f.is_contiguous():
                # the 1x1 conv reads the NCHW feature map as it is (pixel-contiguous A operand): no transposed copy
                y = ops.linear(f, c["w"][l], c["b"][l], w_transposed=True, x_nchw=True)
            else:
                tok = ops.nchw_to_tokens(f)
                y = ops.linear(tok, c["w"][l], c["b"][l], w_transposed=fast, impl=L.IMPL_AUTO if fast else L.IMPL_SIMT)
            ops.groupnorm_tokens_into(y, c["gw"][l], c["gb"][l], src, off, groups=32)
            off += shapes[l][0] * shapes[l][1]
        return src, shapes, c

    _emrt_train = False

    def train(self, mode: bool = True):
        """`.train()` (train.py:139) selects the differentiable path, `.eval()` the fused inference kernels; a freshly
        constructed model runs the inference path."""
        self._emrt_train = bool(mode)
        return super().train(mode)

    def forward(self, src_feats: Sequence[torch.Tensor], src_psp, src_mask=None):
        if src_mask is not None:
            if self._emrt_train and torch.is_grad_enabled():
                raise L.EmrtError("the masked path (src_mask) is built for inference; EMRT never passes a mask (paddle_EMRT.py:265)")
            with torch.no_grad():
                return self._forward_eval(src_feats, src_psp, src_mask)
        if self._emrt_train and torch.is_grad_enabled() and (any(t.requires_grad for t in list(src_feats) + [src_psp])
                                        or any(p.requires_grad for p in self.parameters())):
            return self._forward_train(src_feats, src_psp)
        with torch.no_grad():
            return self._forward_eval(src_feats, src_psp)

    def _forward_train(self, src_feats: Sequence[torch.Tensor], src_psp):
        """t_e_d.py:416-473 with gradients to every parameter the reference trains (input_proj, level_embed, encoder,
        query_pos_embed, reference_points, decoder; tgt_embed is unused, :368) and to the inputs: a composition of
        kernel-backed autograd Functions (emrt_b200/autograd.py)."""
        from . import autograd as A
        x0 = src_feats[0]
        B, dtype, dev = x0.shape[0], x0.dtype, x0.device
        shapes = tuple((int(f.shape[2]), int(f.shape[3])) for f in src_feats)
        impl = L.IMPL_AUTO if dtype == torch.bfloat16 else L.IMPL_SIMT
        srcs = []
        for l, f in enumerate(src_feats):                                                # input_proj (:417-419)
            conv, gn = getattr(self.input_proj[l], "0"), getattr(self.input_proj[l], "1")
            w_kn = conv.weight.reshape(conv.weight.shape[0], -1).t()                      # [Cin, C] = Linear [in, out]
            y = A.linear(A.TokensFn.apply(f), w_kn, conv.bias, impl=impl)
            srcs.append(A.GroupNormTokensFn.apply(y, gn.weight, gn.bias, 1e-5))
        src = torch.cat(srcs, 1)
        pos = A.PosEmbedFn.apply(self.level_embed.weight, self._sine(shapes, dev), shapes, dtype)
        mask = torch.ones((B, src.shape[1]), dtype=torch.float32, device=dev)
        memory = self.encoder(src, shapes, mask, pos)
        ref_dec = A.ReferencePointsFn.apply(self.query_pos_embed.weight, self.reference_points.weight,
                                            self.reference_points.bias, len(shapes))
        qpos = self.query_pos_embed.weight[None].to(dtype)
        tgt = A.TokensFn.apply(src_psp)
        hs = self.decoder(tgt, memory, ref_dec, shapes, mask, qpos)
        return hs, memory

    def _sine(self, shapes, device):
        """Sine position embedding of every level, fp32 [Lv, C] on the device (input-independent; cached per shape)."""
        key = (shapes, str(device))
        hit = getattr(self, "_sine_cache", None)
        if hit is None or hit[0] != key:
            sine = np.concatenate([position_embedding_sine_host(h, w, self.hidden_dim // 2) for (h, w) in shapes], 0)
            self._sine_cache = (key, torch.from_numpy(sine).to(device))
        return self._sine_cache[1]

    def _masked_constants(self, src_mask, shapes, device, dtype):
        """The input-dependent tensors of the masked path (t_e_d.py:408-415,440-451,466-467), from the padding mask alone:
        per-level nearest-resized masks -> mask_flatten [B, Lv], valid ratios [B, L, 2], the masked sine position embedding
        + level embedding [B, Lv, C].  Index / table work on a [B, H, W] mask: done on the host in the reference's float32
        operation order (one device -> host read of the mask); the activations never leave the device."""
        m = src_mask.detach().to("cpu", torch.float32) != 0
        B, H, W = m.shape
        C_ = self.hidden_dim
        npf = C_ // 2
        lvl = self.level_embed.weight.detach().float().cpu()
        masks, pos, vr = [], [], []
        # host torch ops in the reference's own order (position_encoding.py:60-75).  Padded columns / rows make the normalised
        # coordinate (0 - 0.5) / (0 + 1e-6) * 2 pi ~ -3e6: sin / cos of that is ill-conditioned in float32, so the values at
        # PADDED tokens depend on the library's range reduction (they are garbage positions in the reference too).
        dim_t = 2 * (torch.arange(npf) // 2).to(torch.float32)
        dim_t = 10000.0 ** (dim_t / npf)
        for l, (h, w) in enumerate(shapes):
            iy = torch.floor(torch.arange(h) * (H / h)).long()               # F.interpolate(mode='nearest'), :440
            ix = torch.floor(torch.arange(w) * (W / w)).long()
            ml = m[:, iy][:, :, ix].to(torch.float32)                        # [B, h, w]
            vr.append(torch.stack([ml[:, 0, :].sum(1) / w, ml[:, :, 0].sum(1) / h], -1))          # (w, h) order, :408-415
            y_embed = ml.cumsum(1, dtype=torch.float32)
            x_embed = ml.cumsum(2, dtype=torch.float32)
            y_embed = (y_embed + -0.5) / (y_embed[:, -1:, :] + 1e-6) * (2 * math.pi)
            x_embed = (x_embed + -0.5) / (x_embed[:, :, -1:] + 1e-6) * (2 * math.pi)
            px, py = x_embed.unsqueeze(-1) / dim_t, y_embed.unsqueeze(-1) / dim_t
            px = torch.stack((px[:, :, :, 0::2].sin(), px[:, :, :, 1::2].cos()), dim=4).flatten(3)
            py = torch.stack((py[:, :, :, 0::2].sin(), py[:, :, :, 1::2].cos()), dim=4).flatten(3)
            pos.append(torch.cat((py, px), dim=3).reshape(B, h * w, -1) + lvl[l].reshape(1, 1, -1))
            masks.append(ml.reshape(B, h * w))
        to = lambda t: t.contiguous().to(device)
        return (to(torch.cat(masks, 1)), to(torch.stack(vr, 1)), to(torch.cat(pos, 1)).to(dtype).contiguous())

    def _forward_eval(self, src_feats: Sequence[torch.Tensor], src_psp, src_mask=None):
        src, shapes, c = self.project_inputs(src_feats)
        B, Lv, dev = src.shape[0], src.shape[1], src.device
        tgt = ops.nchw_to_tokens(src_psp)                                              # src_psp.transpose([0, 2, 1]) (:469)
        if src_mask is None:
            mask = torch.ones((B, Lv), dtype=torch.float32, device=dev)                # mask_flatten (:451)
            memory = self.encoder(src, shapes, mask, c["pos"])
            hs = self.decoder(tgt, memory, c["ref_dec"], shapes, mask, c["qpos"])
            return hs, memory
        # masked path (:440-447,466-467): per-image valid ratios scale the reference points, the position embedding follows
        # the mask, padded value pixels are zeroed inside MSDeformableAttention (value_mask)
        mask, vr, pos = self._masked_constants(src_mask, shapes, dev, src.dtype)
        memory = self.encoder(src, shapes, mask, pos, vr)
        ref_dec = (c["ref_dec"][:, :, :1, :] * vr[:, None]).contiguous()               # rp.unsqueeze(2) * valid_ratios.unsqueeze(1)
        hs = self.decoder(tgt, memory, ref_dec, shapes, mask, c["qpos"])
        return hs, memory
