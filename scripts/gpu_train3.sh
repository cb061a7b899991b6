#!/bin/bash
TAG=${1:-t3}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_msda.py -m gpu -q --timeout 300 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest.log
echo "== eager"; timeout 600 python scripts/bench_train.py --steps 10 2>&1 | tail -1 | tee gpurun_out/${TAG}_train_eager.json
echo "== graph"; timeout 600 python scripts/bench_train.py --steps 10 --graph 2>&1 | tail -1 | tee gpurun_out/${TAG}_train_graph.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_train_launches.csv \
    python scripts/bench_train.py --steps 2 --warmup 1 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_train_launches.csv > gpurun_out/${TAG}_train_launches.md; head -14 gpurun_out/${TAG}_train_launches.md; tail -1 gpurun_out/${TAG}_train_launches.md
