#!/bin/bash
TAG=${1:-d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decoder.py -m gpu -q --timeout 300 2>&1 | tail -40 > gpurun_out/${TAG}_pytest.log; tail -36 gpurun_out/${TAG}_pytest.log
