#!/bin/bash
TAG=${1:-b2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_msda.py -m gpu -q -x --timeout 300 -k "bwd or backward" 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest.log
echo "== training step, windowed backward"
timeout 600 python scripts/bench_train.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_train.json
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:msda_gather_bwd_win -s 1 -c 1 -o gpurun_out/${TAG}_bwd \
    python scripts/bench_train.py --steps 1 --warmup 1 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${TAG}_train_launches.csv \
    python scripts/bench_train.py --steps 2 --warmup 1 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_train_launches.csv > gpurun_out/${TAG}_train_launches.md; head -6 gpurun_out/${TAG}_train_launches.md
