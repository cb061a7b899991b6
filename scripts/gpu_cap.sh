#!/bin/bash
TAG=${1:-cap}; K=${2:-mha_small}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:$K -s 1 -c 1 -o gpurun_out/${TAG}_$K \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/${TAG}_$K.ncu-rep
