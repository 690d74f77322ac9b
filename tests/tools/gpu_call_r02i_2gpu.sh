#!/bin/bash
# two GPUs: phase timing of the sharded big-image assembly, then the bench line at N=2
set -x
mkdir -p gpurun_out
cat > /tmp/bigtime.py <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch, torch.distributed as dist
import stardist_b200 as sd, bench_data
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = sd.Config2D(n_rays=32)
model = sd.StarDist2D(cfg, name=None, basedir=None, weights=bench_data.bench_weights_2d(cfg))
tile, _ = bench_data.synthetic_image((1024, 1024), seed=0)
img = bench_data.TiledImage(tile, (8, 8))
kw = dict(axes='YX', block_size=2304, min_overlap=128, context=96, show_progress=False)
for it in range(3):
    if it == 2: os.environ["STARDIST_B200_BIG_TIMING"] = "1"
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    labels, polys = model.predict_instances_big(img, **kw)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    if rank == 0: print("pass %d: %.3f s" % (it, dt), flush=True)
if world > 1: dist.destroy_process_group()
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 /tmp/bigtime.py 2>&1 | grep -E "big rank|pass" | tail -20
timeout 120 python /tmp/bigtime.py 2>&1 | grep -E "big rank|pass" | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02i_bench_2gpu.json 2> gpurun_out/r02i_bench_2gpu.err; tail -c 300 gpurun_out/r02i_bench_2gpu.err
python - <<'PY'
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r02i_bench_2gpu.json') if l.startswith('{')][-1]
    print({k: d.get(k) for k in ('n_gpus','value','ms_per_step','value_3d','ms_per_step_3d')})
    print(d['big_2d']['seconds'], d['big_2d']['value'], d['big_3d']['seconds'], d['big_3d']['value'])
except Exception as e: print('bench json', e)
PY
