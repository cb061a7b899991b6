#!/bin/bash
# Window-centre hint: parity, then bench with / without the hint at 8 / 10 / 12 warps, then a full capture.
TAG=${1:-g7}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_msda.py tests/test_gpu_edge_cases.py tests/test_gpu_encoder.py -m gpu -q -x --timeout 300 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest.log
for cfg in "10 7 0" "10 7 1" "12 7 0" "8 7 0" "10 6 0" "10 5 0" "12 5 0"; do
  set -- $cfg
  echo "== WARPS=$1 R=$2 NO_HINT=$3"
  if [ "$3" = "1" ]; then export EMRT_WIN_NO_HINT=1; else unset EMRT_WIN_NO_HINT; fi
  EMRT_WIN_WARPS=$1 EMRT_WIN_R=$2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('ms/step', round(d['ms_per_step'],3), 'gather us', round(r['avg_launch_ms']*1e3,1), 'GB/s', round(r['achieved']))"
done 2>&1 | tee gpurun_out/${TAG}_sweep.log
unset EMRT_WIN_NO_HINT
EMRT_WIN_WARPS=10 timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:msda_gather_fwd_win -s 1 -c 1 -o gpurun_out/${TAG}_win_w10 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
