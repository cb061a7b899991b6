#!/bin/bash
# One gpurun call: full ncu captures of the step's top kernels (encoder gather, the three projections, stitch).
# usage: scripts/gpu_prof.sh <tag>
TAG=${1:-prof}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
NCU="timeout 900 ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:msda_gather_fwd -s 1 -c 1 -o gpurun_out/${TAG}_gather $B > gpurun_out/${TAG}_ncu_gather.log 2>&1
$NCU -k regex:linear_tcgen05 -s 0 -c 3 -o gpurun_out/${TAG}_linear $B > gpurun_out/${TAG}_ncu_linear.log 2>&1
$NCU -k regex:stitch_argmax -s 0 -c 1 -o gpurun_out/${TAG}_stitch $B > gpurun_out/${TAG}_ncu_stitch.log 2>&1
for n in 2 4 8; do python bench.py --steps 10 --warmup 3 --no-cpu-baseline --images $n > gpurun_out/${TAG}_bench_img$n.log 2>&1; done
ls -la gpurun_out
