#!/bin/bash
TAG=${1:-q5}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "decoder or configs or reference_pin or paddle or encoder" 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.log 2>&1; python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.log').read().strip().splitlines()[-1]); r=d['roofline']
print('img/s', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'gather us', round(r['avg_launch_ms']*1e3,1))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 -k regex:groupnorm --csv --log-file gpurun_out/${TAG}_gn.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_gn.csv | head -6
