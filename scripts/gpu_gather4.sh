#!/bin/bash
# Occupancy sweep of the window-staged gather: warps per CTA x region size (1 or 2 CTAs per SM by shared memory).
TAG=${1:-g4}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_msda.py -m gpu -q --timeout 300 -k "window" 2>&1 | tail -2
for cfg in "8 8 16" "12 8 16" "16 16 16" "24 16 16" "24 8 16" "16 8 32" "24 8 32"; do
  set -- $cfg
  echo "== WARPS=$1 TH=$2 TW=$3"
  EMRT_WIN_WARPS=$1 EMRT_WIN_TH=$2 EMRT_WIN_TW=$3 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('ms/step', round(d['ms_per_step'],3), 'gather us', round(r['avg_launch_ms']*1e3,1), 'GB/s', round(r['achieved']))"
done 2>&1 | tee gpurun_out/${TAG}_sweep.log
