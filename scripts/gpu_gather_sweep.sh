#!/bin/bash
# Encoder gather time (CUDA events inside the bench step) under the kernel-selection switches:
#   default = compiled-in geometry (msda_gather_win7.cu), dynamic batch claiming
#   EMRT_WIN_STATIC=1  static dealing;  EMRT_WIN_GENERIC=1  the run-time-geometry kernel of round 1;  EMRT_WIN_NO_HINT=1
TAG=${1:-x}
mkdir -p gpurun_out
run() {
  env "$@" timeout 600 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('%-44s step %.3f ms  gather %.4f ms  frac %.3f  on_chip %.3f' % ('$*', d['ms_per_step'], r['avg_launch_ms'], r['frac'], r['on_chip']['frac']))"
}
{ run EMRT_X=0; run EMRT_WIN_STATIC=1; run EMRT_WIN_GENERIC=1; run EMRT_WIN_NO_HINT=1; run EMRT_WIN_GENERIC=1 EMRT_WIN_NO_HINT=1; } | tee gpurun_out/${TAG}_gather_sweep.txt
