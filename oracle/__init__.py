"""CPU oracle for the EMRT hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``emrt_b200/`` may import this package: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs use it, and only as the checker / the timed CPU baseline.

PARITY: the reference (peach-xiao/EMRT) ships no tests, golden vectors or fixtures, and its
arithmetic lives in PaddlePaddle (>=2.1, <2.5), which cannot be installed in this image.  The
restatement is pinned to the reference's OWN SOURCES executed here on a torch mapping of the Paddle
operators (paddle_on_torch.py, run_reference.py -> tests/golden/ref_*.npz, tests/test_reference_pin.py);
Paddle's own kernels remain unverified — see emrt_oracle.py.
"""
from .emrt_oracle import *  # noqa: F401,F403
