#!/bin/bash
TAG=${1:-b4}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_msda.py -m gpu -q -x --timeout 300 -k "bwd or backward" 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest.log
echo "== timing experiment: plain stores instead of shared-memory reductions (wrong results, timing only)"
EMRT_BWD_WIN_TIMING_PLAIN_STORES=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 -k regex:msda_gather_bwd_win --csv --log-file gpurun_out/${TAG}_plain.csv \
    python scripts/bench_train.py --steps 2 --warmup 1 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_plain.csv > gpurun_out/${TAG}_plain.md; head -4 gpurun_out/${TAG}_plain.md
echo "== 2 GPUs: bench + training step"
