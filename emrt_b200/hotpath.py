"""The hot path as one callable.  Default (``mode="full"``): the reference's whole ``EncoderDecoder.forward``
(transformer_encoder_decoder.py:416-473, constructed as paddle_EMRT.py:241-249 does: input_proj on the C3-C5 features,
position / level embedding, 4 encoder layers [conv branch + MSDeformableAttention + LayerNorm + FFN], reference-point head,
2 decoder layers [110-token self-attention + MSDeformableAttention cross-attention + FFN]) followed by the head tail
(x2 upsample + sliding-window stitch + softmax + argmax; paddle_EMRT.py:178-180, src/api/infer.py:69-79,150-154).
``mode="tokens"`` is the earlier definition (token inputs: TransformerEncoder + the 2 bare decoder MSDA calls + head
tail); ``mode="msda"`` the round-1 starting one (4 + 2 bare MSDA calls + head tail)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from . import ops
from .msda import MSDeformableAttention
from . import refpoints
from . import synthetic


class HotPath:
    def __init__(self, device, tile=512, num_classes=7, embed_dim=256, num_heads=8, num_points=6, num_enc=4,
                 num_dec=2, num_queries=110, seed=1234, gemm_impl=L.IMPL_AUTO, full_encoder=True, mode=None):
        self.device = device
        self.mode = mode or ("tokens" if full_encoder else "msda")
        full_encoder = self.mode == "tokens"
        self.tile, self.nc, self.C = tile, num_classes, embed_dim
        self.shapes = synthetic.level_shapes(tile)
        self.Lv = sum(h * w for h, w in self.shapes)
        self.num_queries = num_queries
        self.enc: List[MSDeformableAttention] = []
        self.dec: List[MSDeformableAttention] = []
        self.model = None
        if self.mode == "full":
            from .decoder import EncoderDecoder
            m = EncoderDecoder(hidden_dim=embed_dim, dim_feedforward=1024, backbone_num_channels=[512, 1024, 2048],
                               dropout=0.1, activation="relu", num_feature_levels=3, nhead=num_heads,
                               num_encoder_layers=num_enc, num_decoder_layers=num_dec, num_encoder_points=num_points,
                               num_decoder_points=num_points, nclass=num_classes)
            st = synthetic.encoder_decoder_state(seed, embed_dim, 1024, num_heads, 3, num_points, num_enc, num_dec)
            with torch.no_grad():
                sd = m.state_dict()
                for k in sd:
                    sd[k].copy_(torch.from_numpy(st[k]))
            for mod in m.modules():
                if hasattr(mod, "gemm_impl"):
                    mod.gemm_impl = gemm_impl
            self.model = m.to(device).requires_grad_(False)
            self.encoder = None
            return
        for i in range(num_enc + num_dec):
            m = MSDeformableAttention(embed_dim, num_heads, len(self.shapes), num_points).to(device)
            st = synthetic.msda_state(seed + i, embed_dim, num_heads, len(self.shapes), num_points)
            with torch.no_grad():
                for name, arr in st.items():
                    mod, leaf = name.split(".")
                    getattr(getattr(m, mod), leaf).copy_(torch.from_numpy(arr))
            m.gemm_impl = gemm_impl
            (self.enc if i < num_enc else self.dec).append(m)
        self.encoder = None
        if full_encoder:
            from .encoder import TransformerEncoder, TransformerEncoderLayer
            enc = TransformerEncoder(TransformerEncoderLayer(embed_dim, num_heads, 1024, 0.1, "relu", len(self.shapes),
                                                             num_points), num_enc)
            with torch.no_grad():
                for i, layer in enumerate(enc.layers):
                    st = synthetic.encoder_layer_state(seed + 10 * i, embed_dim, 1024, num_heads, len(self.shapes), num_points)
                    sd = layer.state_dict()
                    for k in sd:
                        sd[k].copy_(torch.from_numpy(st[k]))
                    layer.gemm_impl = gemm_impl
                    layer.self_attn.gemm_impl = gemm_impl
            self.encoder = enc.to(device).requires_grad_(False)
            self.shapes_t = torch.tensor(self.shapes)
        rng = np.random.Generator(np.random.PCG64(seed + 100))
        self.ref_enc = refpoints.get_reference_points(self.shapes, device=device)   # cached + tagged pixel_grid
        ref_dec = rng.uniform(0.05, 0.95, size=(1, num_queries, 1, 2)).astype(np.float32)
        self.ref_dec = torch.from_numpy(np.ascontiguousarray(np.repeat(ref_dec, len(self.shapes), axis=2))).to(device)
        self.gather_events: Optional[list] = None

    def msda_stack(self, src, pos, tgt, qpos, mask=None):
        """src [B,Lv,C], pos [1|B,Lv,C], tgt [B,Nq,C], qpos [1,Nq,C] -> (memory [B,Lv,C], hs [B,Nq,C])."""
        x = src
        if self.encoder is not None:
            x = self.encoder(x, self.shapes, mask, pos)           # reference points cached inside (t_e_d.py:230-239)
        else:
            for m in self.enc:
                q = ops.add_bcast(x, pos)                          # with_pos_embed (t_e_d.py:198)
                x = m(q, self.ref_enc, x, self.shapes, mask)
        t = tgt
        for m in self.dec:
            q = ops.add_bcast(t, qpos)                             # t_e_d.py:288
            t = m(q, self.ref_dec, x, self.shapes, mask)
        return x, t

    def head_tail(self, half_logits, win_img, win_y0, win_x0, n_img, H, W, labels=None):
        return ops.stitch_argmax_fused(half_logits, win_img, win_y0, win_x0, n_img, H, W, label_dtype=torch.uint8,
                                       labels=labels)[0]

    def step(self, batch, labels=None):
        if self.model is not None:
            hs, mem = self.model(batch["feats"], batch["psp"])               # EncoderDecoder.forward(src_feats, src_psp)
            lab = self.head_tail(batch["half_logits"], batch["win_img"], batch["win_y0"], batch["win_x0"],
                                 batch["n_img"], batch["H"], batch["W"], labels)
            return lab, hs[0]
        mem, hs = self.msda_stack(batch["src"], batch["pos"], batch["tgt"], batch["qpos"], batch.get("mask"))
        lab = self.head_tail(batch["half_logits"], batch["win_img"], batch["win_y0"], batch["win_x0"],
                             batch["n_img"], batch["H"], batch["W"], labels)
        return lab, hs
