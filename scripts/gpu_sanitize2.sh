#!/bin/bash
TAG=${1:-san2}
mkdir -p gpurun_out
CS="timeout 900 compute-sanitizer --error-exitcode 9"
$CS --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_msda.py -m gpu -q -x --timeout 800 -k "bwd_windowed and 128 and v2" > gpurun_out/${TAG}_racecheck_bwd.log 2>&1
grep -E "hazard|Write access|Read access|RACECHECK" gpurun_out/${TAG}_racecheck_bwd.log | sed 's/\[[0-9]* hazards\]//' | sort | uniq -c | sort -rn | head -30
$CS --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_msda.py -m gpu -q -x --timeout 800 -k "window_staged_kernel and 128" > gpurun_out/${TAG}_racecheck_fwd.log 2>&1
grep -E "hazard|Write access|Read access|RACECHECK" gpurun_out/${TAG}_racecheck_fwd.log | sed 's/\[[0-9]* hazards\]//' | sort | uniq -c | sort -rn | head -20
