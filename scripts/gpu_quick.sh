#!/bin/bash
# Quick GPU check: parity tests + a short bench + launch list.  usage: scripts/gpu_quick.sh <tag> [pytest -k expr]
TAG=${1:-quick}; KEXPR=${2:-}
mkdir -p gpurun_out
if [ -n "$KEXPR" ]; then
  timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 -k "$KEXPR" 2>&1 | tail -30 > gpurun_out/${TAG}_pytest.log
else
  timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -40 > gpurun_out/${TAG}_pytest.log
fi
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -8 gpurun_out/${TAG}_pytest.log; tail -c 1500 gpurun_out/${TAG}_bench.log
python scripts/launch_summary.py gpurun_out/${TAG}_launches.csv | head -14
