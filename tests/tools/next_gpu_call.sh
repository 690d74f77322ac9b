#!/bin/bash
# First GPU call of the next round: validate and time what this round could only pin on the CPU.
#   gpurun --timeout 600 -- 'bash tests/tools/next_gpu_call.sh'
set -x
mkdir -p gpurun_out
# 1. everything that is on by default (incl. the 12 reference test_nms_accuracy cases added after the last GPU run)
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -5
# 2. S3 / S4 volume stages on pre-normalised planes (bit-identical on the host build): goldens, then timing old vs new
STARDIST_B200_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_gpu_3d.py -m gpu -q -k normalised 2>&1 | tail -3
for v in 0 1; do
  timeout 200 python tests/tools/run_3d_full.py 64 256 256 --cell 64 256 256 --skip-r1 --nms3d-variant $v 2>&1 | grep -E "^R2" | tail -1
done
# 3. host staging copy (torch parallel copy into the pinned buffer): 3-D end-to-end time
timeout 200 python tests/tools/run_3d_full.py 128 512 512 --json gpurun_out/next_3d_full.json 2>&1 | grep -E "network|R1|R2"
# 4. the bench line
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/next_bench.json 2> gpurun_out/next_bench.err; tail -c 300 gpurun_out/next_bench.json
