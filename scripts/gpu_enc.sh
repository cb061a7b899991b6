#!/bin/bash
TAG=${1:-e}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_encoder.py -m gpu -q --timeout 300 2>&1 | tail -12 > gpurun_out/${TAG}_pytest.log; tail -5 gpurun_out/${TAG}_pytest.log
for w in 18 72; do timeout 300 python scripts/bench_encoder.py --windows $w 2>&1 | tail -1; done | tee gpurun_out/${TAG}_enc.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 200 --csv --log-file gpurun_out/${TAG}_enc_launches.csv \
    python scripts/bench_encoder.py --windows 18 --steps 3 --warmup 2 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_enc_launches.csv > gpurun_out/${TAG}_enc_launches.md; head -16 gpurun_out/${TAG}_enc_launches.md
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:conv3x3_tokens_tc -s 2 -c 1 -o gpurun_out/${TAG}_conv \
    python scripts/bench_encoder.py --windows 18 --steps 1 --warmup 1 > /dev/null 2>&1
