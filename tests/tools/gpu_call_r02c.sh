#!/bin/bash
# round 2, third GPU call: full GPU suite WITHOUT -x, 3-D NMS stage counters, bench
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --durations=15 2>&1 | tail -120 > gpurun_out/r02c_pytest.log; grep -E "passed|failed|FAILED|ERROR" gpurun_out/r02c_pytest.log | head -40
timeout 300 python tests/tools/diag_nms3d.py 64 256 256 2>&1 | grep -E "NMS3D|time" | tail -45 > gpurun_out/r02c_diag3d.log; tail -8 gpurun_out/r02c_diag3d.log
timeout 300 python tests/tools/diag_nms3d.py 64 256 256 --aniso 2>&1 | grep -E "NMS3D|time" | tail -8 > gpurun_out/r02c_diag3d_aniso.log; tail -4 gpurun_out/r02c_diag3d_aniso.log
timeout 600 python bench.py --steps 10 --warmup 3 --skip-big --no-cpu-baseline > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; tail -c 300 gpurun_out/r02c_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02c_bench.json'))
print({k: d.get(k) for k in ('value','ms_per_step','value_3d','ms_per_step_3d')})
print(d['config']['stages_ms'], d['config']['nms_kernels_ms'])
print(d['config'].get('stages_ms_3d'), d['config'].get('nms3d_kernels_ms'), d['config'].get('peak_device_memory_gb_3d'))
PY
