#!/bin/bash
TAG=${1:-t}
mkdir -p gpurun_out
timeout 600 python scripts/bench_train.py --steps 5 --warmup 2 2>&1 | tail -2 | tee gpurun_out/${TAG}_train.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/${TAG}_train_launches.csv \
    python scripts/bench_train.py --steps 2 --warmup 1 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_train_launches.csv > gpurun_out/${TAG}_train_launches.md; head -30 gpurun_out/${TAG}_train_launches.md
