"""per-round phase times inside the 2-D NMS tail kernel on the bench image (verbose=2 of sdb_nms2d)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import stardist_b200 as sd, bench_data
from stardist_b200 import _lib
if len(sys.argv) > 1: _lib.load().sdb_nms2d_set_tail(int(sys.argv[1]))
cfg = sd.Config2D(n_rays=32)
model = sd.StarDist2D(cfg, name=None, basedir=None, weights=bench_data.bench_weights_2d(cfg))
img, _ = bench_data.synthetic_image((1024, 1024), seed=0)
for _ in range(3): model.predict_instances(img)
labels, res = model.predict_instances(img, nms_kwargs=dict(verbose=2))
print(len(res['prob']))
