"""CPU oracle for the EMRT hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``emrt_b200/`` may import this package: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs use it, and only as the checker / the timed CPU baseline.

PARITY UNPINNED: the reference (peach-xiao/EMRT) ships no tests, golden vectors or
fixtures, and its arithmetic lives in PaddlePaddle (>=2.1, <2.5), which cannot be
installed in this image.  See oracle/emrt_oracle.py for how the restatement is pinned
instead (two independent formulations + analytic cases).
"""
from .emrt_oracle import *  # noqa: F401,F403
