"""host-side breakdown of one predict_instances call on the bench image (wall clock, synchronised sections)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, bench_data
from stardist_b200 import Config2D, StarDist2D
cfg = Config2D(n_rays=32)
model = StarDist2D(cfg, name=None, basedir=None, weights=bench_data.bench_weights_2d(cfg))
img, _ = bench_data.synthetic_image((1024, 1024), seed=0)
def sync(): torch.cuda.synchronize()
for rep in range(6):
    T = {}
    sync(); t = time.perf_counter()
    x, axes, axes_net, div_by, perm, resizer, n_tiles, grid, grid_dict, channel = model._predict_setup(img, None, None, None)
    T['setup'] = time.perf_counter() - t; t = time.perf_counter()
    xd = model._to_device(x); sync()
    T['to_device'] = time.perf_counter() - t; t = time.perf_counter()
    cand = model._candidates_from_device_input(xd, resizer.point_bounds('YX'), prob_thresh=0.5); sync()
    T['net+candidates'] = time.perf_counter() - t; t = time.perf_counter()
    labels, res = model._instances_from_candidates_device((1024, 1024), cand, nms_thresh=0.4); sync()
    T['nms+labels+to_host'] = time.perf_counter() - t
    sync(); t = time.perf_counter()
    labels, res = model.predict_instances(img, prob_thresh=0.5, nms_thresh=0.4); sync()
    T['predict_instances (whole)'] = time.perf_counter() - t
    if rep >= 3: print({k: round(1e3 * v, 3) for k, v in T.items()}, len(res['prob']))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for rep in range(12):
    sync(); t = time.perf_counter()
    flush.zero_()
    labels, res = model.predict_instances(img, prob_thresh=0.5, nms_thresh=0.4)
    sync(); ts.append(round(1e3 * (time.perf_counter() - t), 2))
print("bench-style e2e loop (with 256 MiB L2 flush), ms per call:", ts)
ts = []
for rep in range(6):
    sync(); t = time.perf_counter()
    flush.zero_(); sync()
    ts.append(round(1e3 * (time.perf_counter() - t), 2))
print("flush alone:", ts)
