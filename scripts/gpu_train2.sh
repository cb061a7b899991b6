#!/bin/bash
TAG=${1:-t2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_msda.py -m gpu -q -x --timeout 300 -k "backward or window" 2>&1 | tail -3
echo "== eager"; timeout 600 python scripts/bench_train.py --steps 10 2>&1 | tail -1 | tee gpurun_out/${TAG}_train_eager.json
echo "== graph"; timeout 600 python scripts/bench_train.py --steps 10 --graph 2>&1 | tail -3 | tee gpurun_out/${TAG}_train_graph.json
