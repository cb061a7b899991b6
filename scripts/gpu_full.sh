#!/bin/bash
# Full check: all GPU tests, smoke, default bench (+ msda-only for continuity), launch list, full capture of the gather at the bench batch
TAG=${1:-f}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log; tail -4 gpurun_out/${TAG}_pytest.log
python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/${TAG}_bench.log 2>&1; tail -c 2500 gpurun_out/${TAG}_bench.log
timeout 600 python bench.py --msda-only --no-cpu-baseline > gpurun_out/${TAG}_bench_msda.log 2>&1; tail -c 600 gpurun_out/${TAG}_bench_msda.log | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches.md; head -22 gpurun_out/${TAG}_launches.md
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:msda_gather_fwd_win -s 1 -c 1 -o gpurun_out/${TAG}_win \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
