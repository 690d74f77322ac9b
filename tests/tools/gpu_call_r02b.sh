#!/bin/bash
# round 2, second GPU call: full GPU suite (incl. parity at size, tail kernel, S3 bound, culling, prep), bench, A/B switches, launch lists
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=12 2>&1 | tail -40 > gpurun_out/r02b_pytest.log; tail -6 gpurun_out/r02b_pytest.log
STARDIST_B200_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_3d.py -m gpu -q -k normalised 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; tail -c 400 gpurun_out/r02b_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02b_bench.json'))
print({k: d.get(k) for k in ('value','ms_per_step','value_3d','ms_per_step_3d')})
print(d['config']['stages_ms'], d['config']['nms_kernels_ms'])
print(d['config'].get('stages_ms_3d'), d['config'].get('nms3d_kernels_ms'), d['config'].get('peak_device_memory_gb_3d'))
print(d.get('big_2d'), d.get('big_3d'))
PY
# A/B: 2-D NMS host rounds, 3-D normalised-planes variant
STARDIST_B200_NMS2D_TAIL=0 timeout 300 python bench.py --steps 10 --warmup 3 --skip-3d --skip-big --no-cpu-baseline > gpurun_out/r02b_bench_notail.json 2>/dev/null
STARDIST_B200_NMS3D_VARIANT=1 timeout 300 python bench.py --steps 3 --warmup 3 --skip-big --no-cpu-baseline > gpurun_out/r02b_bench_v1.json 2>/dev/null
python - <<'PY'
import json
for f in ('gpurun_out/r02b_bench_notail.json','gpurun_out/r02b_bench_v1.json'):
    try:
        d=json.load(open(f)); print(f, d['ms_per_step'], d['config']['stages_ms'], d.get('ms_per_step_3d'), d['config'].get('nms3d_kernels_ms'))
    except Exception as e: print(f, 'ERR', e)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02b_launches_bench2d.csv python bench.py --steps 2 --warmup 1 --skip-3d --skip-big --no-cpu-baseline > gpurun_out/r02b_ncu_bench.log 2>&1; tail -2 gpurun_out/r02b_ncu_bench.log
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_frontier2|k_pairs|k_fast|k_clip|k_tail|k_threshold|k_gather|k_paint|k_precompute" -c 400 --csv --log-file gpurun_out/r02b_nms2d_dram.csv python bench.py --steps 2 --warmup 1 --skip-3d --skip-big --no-cpu-baseline > /dev/null 2>&1
