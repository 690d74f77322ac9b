#!/bin/bash
# eight GPUs: the bench line at N = 8 (weak 2-D / 3-D replicas, strong-scaled big arms)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02n_bench_8gpu.json 2> gpurun_out/r02n_bench_8gpu.err; tail -c 500 gpurun_out/r02n_bench_8gpu.err
python - <<'PY'
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r02n_bench_8gpu.json') if l.startswith('{')][-1]
    print({k: d.get(k) for k in ('n_gpus','value','ms_per_step','value_3d','ms_per_step_3d')})
    print(d['big_2d']['seconds'], d['big_2d']['value'], d['big_3d']['seconds'], d['big_3d']['value'])
    print(d.get('clocks'))
except Exception as e: print('bench json', e)
PY
