#!/bin/bash
# round 2, call ap: final tree — extended randomised parity + compute-sanitizer memcheck of the small cases
set -u
R=r02ap
mkdir -p gpurun_out
timeout 900 python tools/gpu_fuzz.py 900 150 20000 > gpurun_out/${R}_fuzz.json 2> gpurun_out/${R}_fuzz.err; echo "fuzz rc=$?"; cut -c1-700 gpurun_out/${R}_fuzz.json; tail -2 gpurun_out/${R}_fuzz.err | cut -c1-200
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest -q -x -m gpu \
   "tests/test_gpu_parity.py" -k "c1 or c2 or mixed or disconnected or collision or graded or empty or scrambled or slots or c4_small" > gpurun_out/${R}_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/${R}_memcheck.log
tail -5 gpurun_out/${R}_memcheck.log | cut -c1-200
