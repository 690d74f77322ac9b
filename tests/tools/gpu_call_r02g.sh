#!/bin/bash
# round 2, GPU call g: suite, 3-D counters, bench (no CPU legs), 2-D tail A/B
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=5 --timeout 400 --timeout-method thread 2>&1 | tail -100 > gpurun_out/r02g_pytest.log; grep -E "passed|failed|FAILED|ERROR|Timeout" gpurun_out/r02g_pytest.log | head -30
timeout 200 python tests/tools/diag_nms3d.py 64 256 256 2>&1 | grep -E "NMS3D|time" | tail -2 > gpurun_out/r02g_diag3d.log; tail -2 gpurun_out/r02g_diag3d.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; tail -c 300 gpurun_out/r02g_bench.err
for m in 0 2; do STARDIST_B200_NMS2D_TAIL=$m timeout 200 python bench.py --steps 10 --warmup 3 --skip-3d --skip-big --no-cpu-baseline > gpurun_out/r02g_bench_tail$m.json 2>/dev/null; done
python - <<'PY'
import json
for f in ('gpurun_out/r02g_bench.json','gpurun_out/r02g_bench_tail0.json','gpurun_out/r02g_bench_tail2.json'):
    try:
        d=json.load(open(f))
        print(f, {k: d.get(k) for k in ('value','ms_per_step','value_3d','ms_per_step_3d')})
        print('  ', d['config']['stages_ms'], d['config']['nms_kernels_ms'])
        if 'stages_ms_3d' in d['config']: print('  ', d['config'].get('stages_ms_3d'), d['config'].get('nms3d_kernels_ms'))
        if d.get('big_2d'): print('  ', d['big_2d'].get('seconds'), d.get('big_3d', {}).get('seconds'))
    except Exception as e: print(f, 'ERR', e)
PY
