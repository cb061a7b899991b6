#!/bin/bash
TAG=${1:-g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -k "gather or msda or stitch" 2>&1 | tail -5
for cfg in "7 8 16 8" "7 8 16 12" "7 16 8 12" "6 8 16 12"; do
  set -- $cfg
  echo "== R=$1 TH=$2 TW=$3 WARPS=$4"
  EMRT_WIN_R=$1 EMRT_WIN_TH=$2 EMRT_WIN_TW=$3 EMRT_WIN_WARPS=$4 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('ms/step', round(d['ms_per_step'],3), 'gather us', round(r['avg_launch_ms']*1e3,1), 'GB/s', round(r['achieved']))"
done 2>&1 | tee gpurun_out/${TAG}_sweep.log
