#!/bin/bash
# round 2: two GPUs -- sharded predict_instances_big == serial, bench line at N=2 (weak 2-D / 3-D replicas + strong-scaled big arms)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/tools/run_big_dist.py 4096 2>&1 | grep -E "world|sharded|Error|error|assert" | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02f_bench_2gpu.json 2> gpurun_out/r02f_bench_2gpu.err; tail -c 400 gpurun_out/r02f_bench_2gpu.err
python - <<'PY'
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r02f_bench_2gpu.json') if l.startswith('{')][-1]
    print({k: d.get(k) for k in ('n_gpus','value','ms_per_step','value_3d','ms_per_step_3d')})
    print(d.get('big_2d')); print(d.get('big_3d'))
except Exception as e: print('bench json', e)
PY
