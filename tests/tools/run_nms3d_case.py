"""time one golden 3D NMS case on the GPU (verbose), e.g. python tests/tools/run_nms3d_case.py r96_noise02_thr03"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases, torch
from stardist_b200.lib.stardist3d import c_non_max_suppression_inds
name = sys.argv[1] if len(sys.argv) > 1 else "r96_noise02_thr03"
verbose = int(sys.argv[2]) if len(sys.argv) > 2 else 1
d, p, s, rays, thr, shape = cases.nms3d_inputs(name)
v = np.ascontiguousarray(rays.vertices, np.float32); f = np.ascontiguousarray(rays.faces, np.int32)
c_non_max_suppression_inds(d[:64], p[:64], v, f, s[:64], 1, 1, 0, thr)   # warm up
torch.cuda.synchronize(); t0 = time.perf_counter()
keep = c_non_max_suppression_inds(d, p, v, f, s, 1, 1, verbose, thr)
torch.cuda.synchronize(); print(name, "n", len(d), "kept", int(keep.sum()), "time %.3f s" % (time.perf_counter() - t0))
