#!/bin/bash
# First GPU call of a new round: re-validate and re-measure the state the last round left.
#   gpurun --timeout 1800 -- 'bash tests/tools/next_gpu_call.sh'
# (GPU suite with per-test timeouts, the full bench line incl. the CPU legs, stage counters, phase times of k_tail and the
#  ncu launch lists that profiles/r02_summary.md was built from; for N GPUs see gpu_call_r02f_2gpu.sh / gpu_call_r02l_8gpu.sh)
exec bash "$(dirname "$0")/gpu_call_r02j.sh"
