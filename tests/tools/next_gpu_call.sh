#!/bin/bash
# First GPU call of a new round: re-validate and re-measure the state the last round left.
#   gpurun --timeout 1800 -- 'bash tests/tools/next_gpu_call.sh'
# 1. the cases written after round 2's GPU budget was spent (device NMS / rendering for Rays_Cartesian, Octo, Tetra; marked
#    xfail(strict=False) until they have run once): --runxfail turns them into ordinary tests
# 2. the GPU suite, the full bench line incl. the CPU legs, stage counters, phase times of k_tail and the ncu launch lists
#    that profiles/r02_summary.md was built from (gpu_call_r02j.sh; for N GPUs see gpu_call_r02f_2gpu.sh / gpu_call_r02l_8gpu.sh)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zz_other_rays.py -m gpu -q --runxfail --timeout 120 --timeout-method thread > gpurun_out/next_other_rays.log 2>&1
tail -5 gpurun_out/next_other_rays.log
exec bash "$(dirname "$0")/gpu_call_r02j.sh"
