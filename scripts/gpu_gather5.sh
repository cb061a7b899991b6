#!/bin/bash
# Per-level stage A / no query table / 3-D grid rewrite of the window gather: parity tests, then warps x region sweep,
# then full ncu captures of the 8- and 12-warp variants.
TAG=${1:-g5}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_msda.py tests/test_gpu_edge_cases.py -m gpu -q -x --timeout 300 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest.log
for cfg in "8 8 16 7" "10 8 16 7" "12 8 16 7" "12 8 16 6" "16 16 16 7" "24 16 16 7" "12 8 32 7" "16 8 32 7"; do
  set -- $cfg
  echo "== WARPS=$1 TH=$2 TW=$3 R=$4"
  EMRT_WIN_WARPS=$1 EMRT_WIN_TH=$2 EMRT_WIN_TW=$3 EMRT_WIN_R=$4 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('ms/step', round(d['ms_per_step'],3), 'gather us', round(r['avg_launch_ms']*1e3,1), 'GB/s', round(r['achieved']))"
done 2>&1 | tee gpurun_out/${TAG}_sweep.log
for w in 8 12; do
EMRT_WIN_WARPS=$w timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:msda_gather_fwd_win -s 1 -c 1 -o gpurun_out/${TAG}_win_w$w \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
done
ls -la gpurun_out | tail -5
