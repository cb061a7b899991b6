#!/bin/bash
# Sweep of the window-staged gather's scheduling variants inside the full bench step.
TAG=${1:-g3}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -k "gather or msda or encoder" 2>&1 | tail -3
for cfg in "8 0" "8 1" "7 1" "7 0" "12 0"; do
  set -- $cfg
  echo "== WARPS=$1 STATIC=$2"
  EMRT_WIN_WARPS=$1 EMRT_WIN_STATIC=$2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('ms/step', round(d['ms_per_step'],3), 'gather us', round(r['avg_launch_ms']*1e3,1), 'GB/s', round(r['achieved']))"
done 2>&1 | tee gpurun_out/${TAG}_sweep.log
