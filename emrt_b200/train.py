"""Gradient exchange of the training-step configuration (cfg 4).

The reference trains with ``paddle.DataParallel`` (train.py:116-123): one process per GPU, gradients summed over
ranks and divided by nranks inside ``loss.backward()`` (train.py:153), fused by Paddle into ~25 MB buckets.  This is
the only collective on the hot path.  Here: parameters' ``.grad`` tensors are VIEWS into a few persistent flat fp32
buckets, so the all-reduce (NCCL over NVLink / NVSwitch; ``gloo`` in the CPU tests) runs in place with no pack /
unpack copies, one call per bucket, issued on the communication stream as soon as the step's backward is done.
Parameters that never receive a gradient (``tgt_embed``, ``backbone.fc`` in EMRT) simply contribute zeros.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class GradientBuckets:
    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_mb: float = 25.0,
                 process_group: Optional[dist.ProcessGroup] = None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        self.group = process_group
        cap = max(1, int(bucket_mb * (1 << 20) / 4))
        self.buckets: List[torch.Tensor] = []
        plan, cur, cur_n = [], [], 0
        for p in self.params:
            n = p.numel()
            if cur and cur_n + n > cap:
                plan.append(cur)
                cur, cur_n = [], 0
            cur.append(p)
            cur_n += n
        if cur:
            plan.append(cur)
        for group_params in plan:
            total = sum(p.numel() for p in group_params)
            flat = torch.zeros(total, dtype=torch.float32, device=group_params[0].device)
            off = 0
            for p in group_params:
                if p.dtype != torch.float32:
                    raise TypeError("gradient buckets hold fp32 master gradients")
                p.grad = flat[off:off + p.numel()].view_as(p)      # autograd accumulates into the bucket in place
                off += p.numel()
            self.buckets.append(flat)

    @property
    def nbytes(self) -> int:
        return sum(b.numel() * 4 for b in self.buckets)

    def zero(self):
        for b in self.buckets:
            b.zero_()

    def all_reduce(self, async_op: bool = False):
        """Average the buckets over the ranks (sum / world, train.py:153's DataParallel semantics)."""
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return []
        world = dist.get_world_size(self.group)
        backend = dist.get_backend(self.group)
        works = []
        for b in self.buckets:
            if backend == "nccl":
                works.append(dist.all_reduce(b, op=dist.ReduceOp.AVG, group=self.group, async_op=True))
            else:
                works.append(dist.all_reduce(b, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        if async_op and backend == "nccl":
            return works
        for w in works:
            w.wait()
        if backend != "nccl":
            for b in self.buckets:
                b.div_(world)
        return []


def build_train_step(dev, rank: int, batch: int = 16, tile: int = 512, whole_model: Optional[bool] = None):
    """The cfg-4 step on synthetic data: returns ``(step, buckets, what)``; ``step()`` zeroes the gradient
    buckets and runs forward + backward, ``buckets.all_reduce()`` is the exchange (train.py:153).  With ``whole_model``
    (default: when the EncoderDecoder mirror is differentiable) the step is the reference's whole
    ``EncoderDecoder.forward`` on C3-C5 features + backward into all of its parameters; otherwise the 4 encoder + 2 decoder
    MSDeformableAttention modules on token inputs."""
    from . import synthetic
    from .msda import MSDeformableAttention
    from .refpoints import get_reference_points
    shapes = synthetic.level_shapes(tile)
    Lv = sum(h * w for h, w in shapes)
    B, C, Nq = batch, 256, 110
    g = torch.Generator(device="cpu").manual_seed(rank)
    rnd = lambda *s: torch.randn(s, generator=g).bfloat16().to(dev)
    from . import decoder as _dec
    if whole_model is None:
        whole_model = bool(getattr(_dec, "TRAINABLE", False))
    if whole_model:
        from .decoder import EncoderDecoder
        m = EncoderDecoder(hidden_dim=C, dim_feedforward=1024, backbone_num_channels=[512, 1024, 2048], dropout=0.0,
                           activation="relu", num_feature_levels=3, nhead=8, num_encoder_layers=4, num_decoder_layers=2,
                           num_encoder_points=6, num_decoder_points=6, nclass=7)
        st = synthetic.encoder_decoder_state(1234)
        with torch.no_grad():
            sd = m.state_dict()
            for k in sd:
                sd[k].copy_(torch.from_numpy(st[k]))
        m = m.to(dev).train()
        params = [p for n, p in m.named_parameters() if not n.startswith("tgt_embed")]      # never used (t_e_d.py:368)
        buckets = GradientBuckets(params)
        feats = [(rnd(B, c, tile // s, tile // s) * 0.5).requires_grad_(True) for c, s in zip((512, 1024, 2048), (8, 16, 32))]
        psp = (rnd(B, C, Nq) * 0.5).requires_grad_(True)
        d_mem, d_hs = rnd(B, Lv, C), rnd(1, B, Nq, C)

        def step():
            buckets.zero()
            for t in feats + [psp]:
                t.grad = None
            hs, mem = m(feats, psp)
            torch.autograd.backward([mem, hs], [d_mem, d_hs])
        return step, buckets, ("whole EncoderDecoder (input_proj, 4 encoder layers, 2 decoder layers; %d parameters), fwd + bwd"
                               % sum(p.numel() for p in params))
    mods = []
    for i in range(6):
        mod = MSDeformableAttention(C, 8, 3, 6).to(dev)
        with torch.no_grad():
            for name, arr in synthetic.msda_state(1234 + i).items():
                sub, leaf = name.split(".")
                getattr(getattr(mod, sub), leaf).copy_(torch.from_numpy(arr))
        mods.append(mod)
    buckets = GradientBuckets([p for mod in mods for p in mod.parameters()])
    src, tgt, pos, qpos = rnd(B, Lv, C), rnd(B, Nq, C), rnd(1, Lv, C), rnd(1, Nq, C)
    d_mem, d_hs = rnd(B, Lv, C), rnd(B, Nq, C)
    ref_enc = get_reference_points(shapes, device=dev)
    ref_dec = torch.rand((1, Nq, 1, 2), generator=g).expand(-1, -1, 3, -1).contiguous().to(dev)

    def step():
        buckets.zero()
        mem = src.clone().requires_grad_(True)
        for mod in mods[:4]:
            mem = mod(mem + pos, ref_enc, mem, shapes)      # with_pos_embed: torch add (autograd glue)
        hs = tgt.clone().requires_grad_(True)
        for mod in mods[4:]:
            hs = mod(hs + qpos, ref_dec, mem, shapes)
        torch.autograd.backward([mem, hs], [d_mem, d_hs])
    return step, buckets, "the 4 encoder + 2 decoder MSDeformableAttention modules on token inputs, fwd + bwd"
