#!/bin/bash
# usage: scripts/gpu_multi2.sh <tag> <ngpus>: N-GPU bench (ours + reference arm) and N-GPU training step under torchrun
TAG=${1:-m}; N=${2:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench$N.log 2>&1; tail -1 gpurun_out/${TAG}_bench$N.log | cut -c1-900
timeout 600 $TR --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/${TAG}_ref$N.log 2>&1; tail -1 gpurun_out/${TAG}_ref$N.log | cut -c1-300
timeout 600 $TR --master-port 29513 scripts/bench_train.py --steps 10 > gpurun_out/${TAG}_train$N.log 2>&1; tail -1 gpurun_out/${TAG}_train$N.log
