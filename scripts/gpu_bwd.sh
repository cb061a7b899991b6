#!/bin/bash
# Windowed backward gather: parity tests, memcheck of one small case, training-step timing (windowed vs generic).
TAG=${1:-b1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_msda.py -m gpu -q -x --timeout 300 -k "bwd or backward" 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_msda.py -m gpu -q -x --timeout 500 -k "bwd_windowed and 128" 2>&1 | tail -12 | tee gpurun_out/${TAG}_memcheck.log
echo "== training step, windowed backward"
timeout 600 python scripts/bench_train.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_train.json
echo "== training step, generic backward"
EMRT_GATHER_NO_WIN=1 timeout 600 python scripts/bench_train.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_train_generic.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${TAG}_train_launches.csv \
    python scripts/bench_train.py --steps 2 --warmup 1 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_train_launches.csv > gpurun_out/${TAG}_train_launches.md; head -12 gpurun_out/${TAG}_train_launches.md
