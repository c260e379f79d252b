#!/bin/bash
# round 2, call as: ncu --set full of the C5 kernels (256^3 block, spheres + ellipsoids: refinement + queue partition variants)
set -u
R=r02as
mkdir -p gpurun_out
timeout 700 ncu --set full --import-source on --clock-control none --kernel-name regex:"^(k_classify4|k_heavy_box|k_final|k_connectivity)$" --launch-skip 8 --launch-count 4 \
   -o gpurun_out/${R} python bench.py --workload c5 --cells-per-side 256 --steps 2 --warmup 2 --no-cpu --no-e2e --no-check > gpurun_out/${R}_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/${R}_ncu.log | cut -c1-200; ls -la gpurun_out/${R}.ncu-rep
