#!/bin/bash
TAG=${1:-b5}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_msda.py -m gpu -q -x --timeout 300 -k "bwd or backward" 2>&1 | tail -12 | tee gpurun_out/${TAG}_pytest.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_msda.py -m gpu -q -x --timeout 500 -k "bwd_windowed and 128" 2>&1 | tail -6 | tee gpurun_out/${TAG}_memcheck.log
for w in 16 12; do
echo "== training step, windowed backward v2, $w warps"
EMRT_BWD_WIN_WARPS=$w timeout 600 python scripts/bench_train.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_train_w$w.json
EMRT_BWD_WIN_WARPS=$w timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 -k regex:msda_gather --csv --log-file gpurun_out/${TAG}_train_launches_w$w.csv \
    python scripts/bench_train.py --steps 2 --warmup 1 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_train_launches_w$w.csv > gpurun_out/${TAG}_l_w$w.md; head -4 gpurun_out/${TAG}_l_w$w.md
done
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:msda_gather_bwd_win2 -s 1 -c 1 -o gpurun_out/${TAG}_bwd2 \
    python scripts/bench_train.py --steps 1 --warmup 1 > /dev/null 2>&1
