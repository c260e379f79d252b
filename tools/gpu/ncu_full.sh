# usage: bash tools/gpu/ncu_full.sh <tag> <kernel regex> [extra bench args]   (env NCU_SKIP, NCU_COUNT)
set -u
R=$1; K=$2; shift 2
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name regex:"$K" --launch-skip ${NCU_SKIP:-2} --launch-count ${NCU_COUNT:-1} \
   -o gpurun_out/${R} python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e --no-check "$@" > gpurun_out/${R}_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/${R}_ncu.log | cut -c1-300; ls -la gpurun_out/${R}.ncu-rep
