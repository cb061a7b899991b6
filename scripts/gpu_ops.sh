#!/bin/bash
# Times the four big GEMMs of an encoder layer at the bench geometry (20 back-to-back launches each): bash scripts/gpu_ops.sh TAG
TAG=${1:-x}
mkdir -p gpurun_out
for op in ${EMRT_OPS:-value qproj ffn1 ffn2}; do EMRT_OP_ITERS=20 python scripts/op_once.py $op; done 2>&1 | grep "avg ms" | tee gpurun_out/${TAG}_ops.txt
