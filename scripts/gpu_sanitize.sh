#!/bin/bash
# compute-sanitizer passes over the window kernels (small cases): memcheck forward + backward, racecheck backward v2 / forward
TAG=${1:-san}
mkdir -p gpurun_out
CS="timeout 900 compute-sanitizer --error-exitcode 9"
$CS --tool memcheck python -m pytest tests/test_gpu_msda.py -m gpu -q -x --timeout 800 -k "window_staged_kernel and 128" 2>&1 | tail -4 | tee gpurun_out/${TAG}_memcheck_fwd.log
$CS --tool racecheck python -m pytest tests/test_gpu_msda.py -m gpu -q -x --timeout 800 -k "bwd_windowed and 128 and v2" 2>&1 | tail -6 | tee gpurun_out/${TAG}_racecheck_bwd.log
$CS --tool racecheck python -m pytest tests/test_gpu_msda.py -m gpu -q -x --timeout 800 -k "window_staged_kernel and 128 and float16" 2>&1 | tail -6 | tee gpurun_out/${TAG}_racecheck_fwd.log
