"""torchrun --nproc-per-node N tests/tools/run_big_dist.py [size]
Multi-GPU predict_instances_big (blocks sharded over ranks, NCCL) vs the single-process result on rank 0."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import stardist_b200 as sd, bench_data

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
cfg = sd.Config2D(n_rays=32)
model = sd.StarDist2D(cfg, name=None, basedir=None, weights=bench_data.bench_weights_2d(cfg))
img, _ = bench_data.synthetic_image((size, size), seed=5)
kw = dict(axes='YX', block_size=1024, min_overlap=128, context=96, show_progress=False)
model.predict_instances_big(img[:1200, :1200], **kw)          # warm-up (sharded as well)
if world > 1: dist.barrier()
torch.cuda.synchronize(); t0 = time.perf_counter()
labels, polys = model.predict_instances_big(img, **kw)
torch.cuda.synchronize()
if world > 1: dist.barrier()
dt = time.perf_counter() - t0
if rank == 0:
    n = len(polys['prob'])
    print("world %d: %dx%d -> %d instances in %.3f s (%.0f instances/s)" % (world, size, size, n, dt, n / dt))
    if world > 1:
        # single-process reference on rank 0 (group=None would still see the initialised process group, so
        # run the serial loop explicitly)
        from stardist_b200 import parallel_big
        orig = parallel_big.rank_world
        parallel_big.rank_world = lambda group=None: (0, 1)
        l1, p1 = model.predict_instances_big(img, **kw)
        parallel_big.rank_world = orig
        assert np.array_equal(labels, l1), "sharded label map differs from the serial one"
        assert np.array_equal(polys['points'], p1['points']) and np.array_equal(polys['prob'], p1['prob'])
        print("sharded == serial: OK")
if world > 1:
    dist.barrier(); dist.destroy_process_group()
