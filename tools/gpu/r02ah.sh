#!/bin/bash
# round 2, call ah (2 GPUs): library comm test with the split step
set -u
R=r02ah
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_zx_gpu_comm.py -m gpu -q -x > gpurun_out/${R}_pytest_comm.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${R}_pytest_comm.log
tail -25 gpurun_out/${R}_pytest_comm.log | cut -c1-300
