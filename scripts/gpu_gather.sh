#!/bin/bash
# Gather / GEMM iteration run: parity tests for the kernels, window-kernel parameter sweep, bench + launch list.
TAG=${1:-g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -40 > gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
for cfg in "7 8 16 1" "7 8 16 0" "6 8 16 1" "7 16 8 1" "5 8 16 1"; do
  set -- $cfg
  echo "== R=$1 TH=$2 TW=$3 REC64=$4"
  EMRT_WIN_R=$1 EMRT_WIN_TH=$2 EMRT_WIN_TW=$3 EMRT_WIN_REC64=$4 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('ms/step', round(d['ms_per_step'],3), 'gather us', round(r['avg_launch_ms']*1e3,1), 'GB/s', round(r['achieved']))"
done 2>&1 | tee gpurun_out/${TAG}_sweep.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches.md; head -12 gpurun_out/${TAG}_launches.md
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:msda_gather_fwd_win -s 1 -c 1 -o gpurun_out/${TAG}_win \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_win.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:stitch_argmax -s 0 -c 1 -o gpurun_out/${TAG}_stitch \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_stitch.log 2>&1
