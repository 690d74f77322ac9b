#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_3d.py tests/test_gpu_at_size.py tests/test_gpu_big.py -m gpu -q --durations=5 --timeout 400 --timeout-method thread 2>&1 | tail -80 > gpurun_out/r02h_pytest.log; grep -E "passed|failed|FAILED|ERROR|Timeout|illegal" gpurun_out/r02h_pytest.log | head -30
timeout 120 python tests/tools/diag_nms2d_tail.py 1 2>&1 | grep -E "round|tail|NMS2D" | head -30 > gpurun_out/r02h_tail_phases.log; cat gpurun_out/r02h_tail_phases.log
