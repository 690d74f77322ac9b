#!/bin/bash
# final validation: GPU suite + bench line with the last code state
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=5 --timeout 400 --timeout-method thread 2>&1 | tail -60 > gpurun_out/r02m_pytest.log; grep -E "passed|failed|FAILED|ERROR|Timeout|illegal" gpurun_out/r02m_pytest.log | head -20
timeout 200 python tests/tools/diag_nms3d.py 64 256 256 2>&1 | grep -E "NMS3D|time" | tail -2 > gpurun_out/r02m_diag3d.log; tail -2 gpurun_out/r02m_diag3d.log
timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err; tail -c 300 gpurun_out/r02m_bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02m_bench.json'))
    print({k: d.get(k) for k in ('value','ms_per_step','value_3d','ms_per_step_3d')}, d.get('clocks'))
    print(d['config'].get('stages_ms_3d'), d['config'].get('nms3d_kernels_ms'))
    print(d['big_2d']['seconds'], d['big_3d']['seconds'])
except Exception as e: print('bench json', e)
PY
