#!/bin/bash
TAG=${1:-it}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -k "gather or msda or configs or encoder" 2>&1 | tail -6
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.log 2>&1; python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.log').read().strip().splitlines()[-1]); r=d['roofline']
print('img/s', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'gather us', round(r['avg_launch_ms']*1e3,1), d['clocks'])
PY
