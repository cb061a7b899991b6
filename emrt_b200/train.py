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
