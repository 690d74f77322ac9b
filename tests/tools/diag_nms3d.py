"""stage counters of the 3-D NMS on the bench workload (a 64x256x256 cell): verbose output of sdb_nms3d"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import stardist_b200 as sd, bench_data
shape = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (64, 256, 256)
aniso = (2, 1, 1) if "--aniso" in sys.argv else None
cfg = bench_data.bench_config_3d(96, anisotropy=aniso)
model = sd.StarDist3D(cfg, name=None, basedir=None, weights=bench_data.bench_weights_3d(cfg))
vol, _ = bench_data.synthetic_volume(shape, seed=0, cell=tuple(min(a, b) for a, b in zip(bench_data.CELL_3D, shape)))
model.predict_instances(vol, prob_thresh=0.7, nms_thresh=0.3)
torch.cuda.synchronize(); t0 = time.perf_counter()
labels, res = model.predict_instances(vol, prob_thresh=0.7, nms_thresh=0.3, nms_kwargs=dict(verbose=True))
torch.cuda.synchronize(); print("time %.3f s, %d instances, candidates %d" % (time.perf_counter() - t0, len(res['prob']), model._last_n_cand))
