"""time the 2D NMS of the benchmark image standalone (verbose round statistics)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, bench_data
from stardist_b200 import Config2D, StarDist2D, _lib as L
cfg = Config2D(n_rays=32)
model = StarDist2D(cfg, name=None, basedir=None, weights=bench_data.bench_weights_2d(cfg))
img, _ = bench_data.synthetic_image((1024, 1024), seed=0)
cand = model._predict_sparse_device(img, prob_thresh=0.5)
print("candidates", cand['n'])
lib = L.load()
keep = torch.zeros(cand['n'], dtype=torch.uint8, device='cuda')
ref = None
for mode, verbose in ((0, 0), (0, 2), (0, 0), (1, 0), (1, 2), (1, 0), (1, 0), (2, 1)):
    L.nms2d_set_filter(mode)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    L.check(lib.sdb_nms2d(L.ptr(cand['dist']), L.ptr(cand['points_f32']), cand['n'], 32, 0.4, 1, 1, verbose, L.ptr(keep), L.stream_ptr()))
    torch.cuda.synchronize(); print("filter mode %d: nms2d %.2f ms kept %d" % (mode, 1e3 * (time.perf_counter() - t0), int(keep.sum())))
    if ref is None: ref = keep.clone()
    assert torch.equal(ref, keep), "filter mode changes the result"
print(L.nms2d_filter_stats())
