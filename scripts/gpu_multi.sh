#!/bin/bash
# usage: scripts/gpu_multi.sh <tag> <ngpus>: config tests, 1-GPU bench, N-GPU bench under torchrun
TAG=${1:-m}; N=${2:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "configs" 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log; tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench1.log 2>&1; tail -c 1200 gpurun_out/${TAG}_bench1.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench$N.log 2>&1; tail -c 1500 gpurun_out/${TAG}_bench$N.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_ref.log 2>&1; tail -c 600 gpurun_out/${TAG}_ref.log
