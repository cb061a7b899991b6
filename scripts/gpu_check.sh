#!/bin/bash
# One gpurun call: GPU parity tests, smoke, a short bench, an ncu launch list and one full capture of the gather.
# usage: scripts/gpu_check.sh <tag> [bench flags...]
TAG=${1:-run}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -40 > gpurun_out/${TAG}_pytest.log
python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -60 > gpurun_out/${TAG}_pytest_all.log
python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1
python bench.py --steps 10 --warmup 3 "$@" > gpurun_out/${TAG}_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_gather_fwd -s 4 -c 1 -f -o gpurun_out/${TAG}_gather \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/${TAG}_ncu_gather.log 2>&1
tail -5 gpurun_out/${TAG}_pytest_all.log; cat gpurun_out/${TAG}_smoke.log | tail -3; tail -2 gpurun_out/${TAG}_bench.log
