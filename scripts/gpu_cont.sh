#!/bin/bash
mkdir -p gpurun_out
for f in "--msda-only" "--tokens"; do
timeout 600 python bench.py $f --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$f', 'img/s', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'gather us', round(r['avg_launch_ms']*1e3,1), 'frac', round(r['frac'],3), d['config'].get('windows_per_gpu'))"
done | tee gpurun_out/cont.log
