#!/bin/bash
# Full check of the current tree: all GPU tests, smoke, default bench + reference arm, launch list, full capture of the gather.
TAG=${1:-s}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log; tail -4 gpurun_out/${TAG}_pytest.log
python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/${TAG}_bench.log 2>&1; tail -c 3000 gpurun_out/${TAG}_bench.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.log 2>&1; tail -c 800 gpurun_out/${TAG}_bench_ref.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches.md; head -22 gpurun_out/${TAG}_launches.md
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:msda_gather_fwd_win -s 1 -c 1 -o gpurun_out/${TAG}_win \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
for cfg in "24 16 16" "16 16 16"; do
  set -- $cfg
  echo "== WARPS=$1 TH=$2 TW=$3"
  EMRT_WIN_WARPS=$1 EMRT_WIN_TH=$2 EMRT_WIN_TW=$3 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('ms/step', round(d['ms_per_step'],3), 'gather us', round(r['avg_launch_ms']*1e3,1), 'GB/s', round(r['achieved']))"
done 2>&1 | tee gpurun_out/${TAG}_sweep2.log
