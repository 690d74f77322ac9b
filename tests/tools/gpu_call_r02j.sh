#!/bin/bash
# round 2, final 1-GPU call: GPU suite, bench line (with CPU legs), launch lists 2-D / 3-D, DRAM capture of the 2-D NMS kernels
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=5 --timeout 400 --timeout-method thread 2>&1 | tail -100 > gpurun_out/r02j_pytest.log; grep -E "passed|failed|FAILED|ERROR|Timeout|illegal" gpurun_out/r02j_pytest.log | head -30
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; tail -c 300 gpurun_out/r02j_bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02j_bench.json'))
    print({k: d.get(k) for k in ('value','ms_per_step','value_3d','ms_per_step_3d')}, d['e2e']['value'], d['e2e_3d']['value'])
    print(d['config']['stages_ms'], d['config']['nms_kernels_ms'])
    print(d['config'].get('stages_ms_3d'), d['config'].get('nms3d_kernels_ms'), d['config'].get('peak_device_memory_gb_3d'))
    print(d['big_2d']['seconds'], d['big_3d']['seconds'])
    print(d['cpu_baseline']['value'], d['cpu_baseline']['value_3d'], d['cpu_baseline']['threads_sweep_instances_per_s'])
    print(d['roofline']['frac'], d['roofline_3d']['frac'], d['roofline_other']['nms_labels_2d']['frac'])
except Exception as e: print('bench json', e)
PY
timeout 200 python tests/tools/diag_nms3d.py 64 256 256 2>&1 | grep -E "NMS3D|time" | tail -2 > gpurun_out/r02j_diag3d.log; tail -2 gpurun_out/r02j_diag3d.log
timeout 120 python tests/tools/diag_nms2d_tail.py 1 2>&1 | grep -E "round|tail|NMS2D" | head -20 > gpurun_out/r02j_tail_phases.log; tail -14 gpurun_out/r02j_tail_phases.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02j_launches_bench2d.csv python bench.py --steps 2 --warmup 1 --skip-3d --skip-big --no-cpu-baseline > gpurun_out/r02j_ncu_bench.log 2>&1; tail -1 gpurun_out/r02j_ncu_bench.log | cut -c1-200
cat > /tmp/step3d.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import torch, stardist_b200 as sd, bench_data
cfg = bench_data.bench_config_3d(96)
model = sd.StarDist3D(cfg, name=None, basedir=None, weights=bench_data.bench_weights_3d(cfg))
vol, _ = bench_data.synthetic_volume((64, 256, 256), seed=0)
for _ in range(2): model.predict_instances(vol, prob_thresh=0.7, nms_thresh=0.3)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02j_launches_3d_64x256x256.csv python /tmp/step3d.py > /dev/null 2>&1; wc -l gpurun_out/r02j_launches_3d_64x256x256.csv
