#!/bin/bash
# round 2, fifth GPU call: GPU suite, 3-D counters, full bench line, launch lists
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=8 --timeout 400 --timeout-method thread 2>&1 | tail -150 > gpurun_out/r02e_pytest.log; grep -E "passed|failed|FAILED|ERROR|Timeout" gpurun_out/r02e_pytest.log | head -40
timeout 200 python tests/tools/diag_nms3d.py 64 256 256 2>&1 | grep -E "NMS3D|time" | tail -3 > gpurun_out/r02e_diag3d.log; tail -2 gpurun_out/r02e_diag3d.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; tail -c 300 gpurun_out/r02e_bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02e_bench.json'))
    print({k: d.get(k) for k in ('value','ms_per_step','value_3d','ms_per_step_3d')}, d['e2e']['value'], d['e2e_3d']['value'])
    print(d['config']['stages_ms'], d['config']['nms_kernels_ms'])
    print(d['config'].get('stages_ms_3d'), d['config'].get('nms3d_kernels_ms'), d['config'].get('peak_device_memory_gb_3d'))
    print(d.get('big_2d'), d.get('big_3d'))
    print(d.get('cpu_baseline'))
except Exception as e: print('bench json', e)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02e_launches_bench2d.csv python bench.py --steps 2 --warmup 1 --skip-3d --skip-big --no-cpu-baseline > gpurun_out/r02e_ncu_bench.log 2>&1; tail -1 gpurun_out/r02e_ncu_bench.log | cut -c1-200
