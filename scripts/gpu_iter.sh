#!/bin/bash
# usage: scripts/gpu_iter.sh <tag> : tests + bench variants
TAG=${1:-it}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log; tail -4 gpurun_out/${TAG}_pytest.log
run() { echo "== $1"; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('img/s', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'gather us', round(r['avg_launch_ms']*1e3,1))"; }
run base A=1
run bn128 EMRT_GEMM_BN128=1
run no_tma_store EMRT_GEMM_NO_TMA_STORE=1
for n in 4 8; do echo "== images $n"; timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --images $n 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('img/s', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'gather us', round(r['avg_launch_ms']*1e3,1))"; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches.md; head -9 gpurun_out/${TAG}_launches.md
