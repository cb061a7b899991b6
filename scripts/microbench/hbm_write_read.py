#!/usr/bin/env python
"""Pure-write, pure-read and copy bandwidth of this GPU's HBM at the FFN1 output size (793 MB): the ceilings a write-heavy
kernel (FFN1: 198 MB read, 793 MB written) should be judged against.  python scripts/microbench/hbm_write_read.py"""
import torch

dev = torch.device("cuda", 0)
n = 387072 * 1024
y = torch.empty(n, dtype=torch.bfloat16, device=dev)
x = torch.randn(n, dtype=torch.float32, device=dev).bfloat16()


def timeit(fn, iters=20):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


mb = n * 2 / 1e6
t = timeit(lambda: y.fill_(1.0))
print(f"fill   {mb:.0f} MB written            {t * 1e3:7.1f} us  {mb / t / 1e3:6.2f} TB/s")
t = timeit(lambda: torch.cuda.memset if False else y.zero_())
print(f"zero_  {mb:.0f} MB written            {t * 1e3:7.1f} us  {mb / t / 1e3:6.2f} TB/s")
t = timeit(lambda: y.copy_(x))
print(f"copy   {mb:.0f} MB read + {mb:.0f} written  {t * 1e3:7.1f} us  {2 * mb / t / 1e3:6.2f} TB/s")
t = timeit(lambda: x.view(torch.int16).max())
print(f"max    {mb:.0f} MB read               {t * 1e3:7.1f} us  {mb / t / 1e3:6.2f} TB/s")
q = x[: n // 4]
t = timeit(lambda: torch.relu(q, out=y[: n // 4]) if False else y[: n // 4].copy_(q))
print(f"copy/4 {mb/4:.0f} MB read + {mb/4:.0f} written  {t * 1e3:7.1f} us  {mb / 2 / t / 1e3:6.2f} TB/s")
