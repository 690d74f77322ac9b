"""3D workload (regime R1 + R2): random-init StarDist3D network pass on a synthetic volume, and NMS + label
rendering on prob/dist maps derived analytically from ground-truth ellipsoids (a random-init net has no usable dist head).
Usage: python tests/tools/run_3d_bench.py D H W [n_rays] [prob_thresh]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import stardist_b200 as sd

D, H, W = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (32, 128, 128)
n_rays = int(sys.argv[4]) if len(sys.argv) > 4 else 96
pthr = float(sys.argv[5]) if len(sys.argv) > 5 else 0.7
rng = np.random.default_rng(0)
rays = sd.Rays_GoldenSpiral(n_rays)
cfg = sd.Config3D(rays=rays)
model = sd.StarDist3D(cfg, name=None, basedir=None)

# ---- R1: network pass
vol = rng.uniform(0, 1, (D, H, W)).astype(np.float32)
x = torch.from_numpy(vol[None, ..., None]).cuda()
for _ in range(2): p, d = model.net.forward(x)
torch.cuda.synchronize(); t0 = time.perf_counter()
p, d = model.net.forward(x); torch.cuda.synchronize()
t_net = time.perf_counter() - t0
from stardist_b200.models.weights import unet_layers
flop = 0
sp = np.array([D, H, W], dtype=np.float64)
for l in unet_layers(cfg):
    if l['kind'] == 'pool': sp = sp / 2
    elif l['kind'] == 'up': sp = sp * 2
    elif l['kind'] == 'conv': flop += 2 * np.prod(sp) * l['cin'] * l['cout'] * 27
print("network %dx%dx%d: %.3f s (%.1f algorithmic TFLOP/s of the 3x3x3 convs, CUDA-core fp32 path)" % (D, H, W, t_net, flop / t_net / 1e12))
del p, d, x

# ---- R2: ground-truth ellipsoids -> prob (1 - normalised radius), dist (ray / ellipsoid intersection)
prob = np.zeros((D, H, W), np.float32); dist = np.full((D, H, W, n_rays), 1e-3, np.float32)
occ = np.zeros((D, H, W), bool)
verts = rays.vertices.astype(np.float64)                       # (z, y, x) directions
n_obj = 0
target = 0.30 * D * H * W
filled = 0
for _ in range(200000):
    if filled >= target: break
    r = rng.uniform(5, 9, 3) * np.array([0.6, 1, 1])
    m = np.ceil(r).astype(int) + 1
    c = np.array([rng.integers(m[0], D - m[0]), rng.integers(m[1], H - m[1]), rng.integers(m[2], W - m[2])])
    sl = tuple(slice(c[i] - m[i], c[i] + m[i] + 1) for i in range(3))
    zz, yy, xx = np.mgrid[-m[0]:m[0] + 1, -m[1]:m[1] + 1, -m[2]:m[2] + 1]
    q = np.stack([zz / r[0], yy / r[1], xx / r[2]], -1)             # offsets in the unit-sphere frame
    rn = np.sqrt((q ** 2).sum(-1))
    inside = rn <= 1
    if occ[sl][inside].any(): continue
    occ[sl] |= inside
    v = verts / r                                                   # ray directions in that frame
    a = (v ** 2).sum(-1)                                            # [R]
    b = 2 * (q[..., None, :] * v).sum(-1)                           # [.., R]
    cc = (rn ** 2 - 1)[..., None]
    t = (-b + np.sqrt(np.maximum(b * b - 4 * a * cc, 0))) / (2 * a)
    pr = np.clip(1 - rn, 0, 1)
    pv = prob[sl]; dv = dist[sl]
    pv[inside] = pr[inside]; dv[inside] = t[inside].astype(np.float32)
    n_obj += 1; filled += inside.sum()
n_cand = int((prob > pthr).sum())
print("ground truth: %d ellipsoids, fill %.2f, candidates(prob > %.2f) = %d" % (n_obj, filled / (D * H * W), pthr, n_cand))
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    labels, res = model._instances_from_prediction((D, H, W), prob, dist, prob_thresh=pthr, nms_thresh=0.3)
    torch.cuda.synchronize(); t_post = time.perf_counter() - t0
    print("post-processing (sort + NMS3D + polyhedron_to_label, host arrays in/out): %.3f s -> %d instances (%d labelled voxels)" %
          (t_post, len(res['prob']), int((labels > 0).sum())))
print("R1 + R2: %.3f s per volume -> %.0f instances/s" % (t_net + t_post, len(res['prob']) / (t_net + t_post)))
