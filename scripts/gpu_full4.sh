#!/bin/bash
# Full check of the current tree: all GPU tests, smoke, default bench + reference arm, launch list + gather capture, training step
TAG=${1:-v}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log; tail -4 gpurun_out/${TAG}_pytest.log
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/${TAG}_bench.log 2>&1; tail -c 400 gpurun_out/${TAG}_bench.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.log 2>&1; tail -c 200 gpurun_out/${TAG}_bench_ref.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches.md; head -14 gpurun_out/${TAG}_launches.md
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:msda_gather_fwd_win -s 1 -c 1 -o gpurun_out/${TAG}_win \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 600 python scripts/bench_train.py --steps 10 2>&1 | tail -1 | tee gpurun_out/${TAG}_train.json
timeout 600 python scripts/bench_train.py --steps 10 --graph 2>&1 | tail -1 | tee gpurun_out/${TAG}_train_graph.json
timeout 600 compute-sanitizer --error-exitcode 9 --tool racecheck python -m pytest tests/test_gpu_msda.py -m gpu -q -x --timeout 500 -k "bwd_windowed and 128 and v2" 2>&1 | tail -3 | tee gpurun_out/${TAG}_racecheck_bwd.log
