#!/bin/bash
# round 2, call ar: host-facade GPU tests incl. the taylor_couette example on its O-grid
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_host_gpu.py tests/test_plugin_surface_cpu.py tests/test_foam_adapter_cpu.py -m gpu -q > gpurun_out/r02ar_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02ar_pytest.log | cut -c1-300
