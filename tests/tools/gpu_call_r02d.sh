#!/bin/bash
# round 2, fourth GPU call: full GPU suite (per-test timeout), 3-D stage counters, bench without the CPU legs
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=15 --timeout 400 --timeout-method thread 2>&1 | tail -150 > gpurun_out/r02d_pytest.log; grep -E "passed|failed|FAILED|ERROR|Timeout" gpurun_out/r02d_pytest.log | head -40
timeout 200 python tests/tools/diag_nms3d.py 64 256 256 2>&1 | grep -E "NMS3D|time" | tail -3 > gpurun_out/r02d_diag3d.log; tail -2 gpurun_out/r02d_diag3d.log
timeout 600 python bench.py --steps 10 --warmup 3 --skip-big --no-cpu-baseline > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; tail -c 300 gpurun_out/r02d_bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02d_bench.json'))
    print({k: d.get(k) for k in ('value','ms_per_step','value_3d','ms_per_step_3d')})
    print(d['config']['stages_ms'], d['config']['nms_kernels_ms'])
    print(d['config'].get('stages_ms_3d'), d['config'].get('nms3d_kernels_ms'), d['config'].get('peak_device_memory_gb_3d'))
except Exception as e: print('bench json', e)
PY
