#!/bin/bash
# First GPU call of a round (one B200):  gpurun --timeout 3000 -- 'bash tools/first_gpu_call.sh'
# = the evidence run of round 2 (tools/gpu/r02final.sh): GPU suite, the contract bench line (C4) + the reference arm, the small
# BASELINE configurations, C5 at 256^3, the aux kernels, the ncu launch list and one --set full capture of the three interact
# kernels.  Everything lands in gpurun_out/; nothing that ran under ncu is a bench value.  Multi-GPU: tools/gpu/r02u8.sh (8 GPUs),
# tools/gpu/r02al.sh N; extended randomised parity: python tools/gpu_fuzz.py.
exec bash "$(dirname "$0")/gpu/r02final.sh"
