#!/bin/bash
TAG=${1:-b6}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_msda.py -m gpu -q -x --timeout 300 -k "bwd or backward" 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest.log
for R in 7 6 5 4; do
echo "== training step, v2, R=$R"
EMRT_BWD_WIN_R=$R timeout 600 python scripts/bench_train.py 2>&1 | tail -1 | cut -c1-200
done 2>&1 | tee gpurun_out/${TAG}_R_sweep.log
