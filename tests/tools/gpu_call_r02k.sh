#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_3d.py tests/test_gpu_at_size.py tests/test_gpu_unet_tc.py -m gpu -q --timeout 400 --timeout-method thread 2>&1 | tail -60 > gpurun_out/r02k_pytest.log; grep -E "passed|failed|FAILED|ERROR|Timeout|illegal" gpurun_out/r02k_pytest.log | head -20
timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; tail -c 300 gpurun_out/r02k_bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02k_bench.json'))
    print({k: d.get(k) for k in ('value','ms_per_step','value_3d','ms_per_step_3d')})
    print(d['config'].get('stages_ms_3d'), d['config'].get('nms3d_kernels_ms'))
    print(d['big_2d']['seconds'], d['big_3d']['seconds'])
except Exception as e: print('bench json', e)
PY
