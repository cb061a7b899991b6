#!/bin/bash
TAG=${1:-b3}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_msda.py -m gpu -q -x --timeout 300 -k "bwd or backward" 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest.log
for w in 8 10; do
echo "== training step, windowed backward, $w warps"
EMRT_BWD_WIN_WARPS=$w timeout 600 python scripts/bench_train.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_train_w$w.json
EMRT_BWD_WIN_WARPS=$w timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 -k regex:msda_gather --csv --log-file gpurun_out/${TAG}_train_launches_w$w.csv \
    python scripts/bench_train.py --steps 2 --warmup 1 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_train_launches_w$w.csv | head -5
done
