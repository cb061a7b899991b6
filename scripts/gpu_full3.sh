#!/bin/bash
# Full check of the current tree: all GPU tests, smoke, default bench + reference arm, launch list, training step.
TAG=${1:-u}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log; tail -4 gpurun_out/${TAG}_pytest.log
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/${TAG}_bench.log 2>&1; tail -c 600 gpurun_out/${TAG}_bench.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.log 2>&1; tail -c 300 gpurun_out/${TAG}_bench_ref.log
timeout 600 python scripts/bench_train.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_train.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${TAG}_train_launches.csv \
    python scripts/bench_train.py --steps 2 --warmup 1 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_train_launches.csv > gpurun_out/${TAG}_train_launches.md; head -8 gpurun_out/${TAG}_train_launches.md
