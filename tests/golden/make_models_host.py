"""tests/golden/models_host.json / models_host.npz: host-side behaviour of the reference's model classes, obtained by importing
stardist/models/model2d.py and model3d.py from /root/reference through _refpkg.load_models() (keras / tensorflow / csbdeep
replaced by stubs; only code that never touches the network is exercised):
  * vars(Config2D(**kw)) / vars(Config3D(**kw)) for a list of keyword sets (model2d.py:123-269, model3d.py:129-311),
  * StarDist2D/3D._axes_div_by (model2d.py:566-577, model3d.py:677-691),
  * StarDist2D._instances_from_prediction without labels (model2d.py:512-563; polygons_to_label needs scikit-image) and
    StarDist3D._instances_from_prediction with labels (model3d.py:589-674), dense / sparse / scale / multi-class.
Run in the build container: OMP_NUM_THREADS=1 python tests/golden/make_models_host.py"""
import json, os, sys, types
os.environ.setdefault("OMP_NUM_THREADS", "1")
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refpkg, cases

CFG2D = [dict(), dict(n_rays=64, grid=(2, 2), n_channel_in=3), dict(n_rays=16, n_classes=3, unet_n_depth=2, unet_n_filter_base=16, net_conv_after_unet=64),
         dict(grid=(4, 2), unet_kernel_size=(5, 5), unet_batch_norm=True, unet_pool=(2, 2), backbone='unet')]
CFG3D = [dict(), dict(n_rays=64, grid=(1, 2, 2), anisotropy=(2, 1, 1)), dict(backbone='resnet', grid=(2, 4, 4), n_classes=2), dict(n_channel_in=2, unet_n_depth=3)]


def jsonable(v):
    if isinstance(v, (tuple, list)): return [jsonable(x) for x in v]
    if isinstance(v, dict): return {str(k): jsonable(x) for k, x in v.items()}
    if isinstance(v, (np.integer,)): return int(v)
    if isinstance(v, (np.floating,)): return float(v)
    if isinstance(v, np.ndarray): return jsonable(v.tolist())
    return v


if __name__ == "__main__":
    m2, m3 = _refpkg.load_models()
    rrays = _refpkg.load("stardist.rays3d")
    meta = {"cfg2d": [], "cfg3d": [], "div_by": {}}
    for kw in CFG2D:
        meta["cfg2d"].append([jsonable(kw), jsonable(vars(m2.Config2D(**kw)))])
    for kw in CFG3D:
        k = dict(kw); n = k.pop("n_rays", 96)
        meta["cfg3d"].append([jsonable(kw), jsonable(vars(m3.Config3D(rays=rrays.Rays_GoldenSpiral(n, anisotropy=k.get("anisotropy")), **k)))])
    for i, kw in enumerate(CFG2D):
        fake = types.SimpleNamespace(config=m2.Config2D(**kw))
        meta["div_by"]["2d/%d" % i] = jsonable(m2.StarDist2D._axes_div_by(fake, "YXC"))
    for i, kw in enumerate(CFG3D):
        k = dict(kw); n = k.pop("n_rays", 96)
        fake = types.SimpleNamespace(config=m3.Config3D(rays=rrays.Rays_GoldenSpiral(n, anisotropy=k.get("anisotropy")), **k))
        meta["div_by"]["3d/%d" % i] = jsonable(m3.StarDist3D._axes_div_by(fake, "ZYXC"))
    json.dump(meta, open(os.path.join(HERE, "models_host.json"), "w"), indent=0, sort_keys=True)

    out = {}
    thr = types.SimpleNamespace(prob=0.9, nms=0.3)
    # ---- 2-D: dense, dense with grid + scale, sparse with multi-class probabilities
    name = "r32_356x299"
    shape, radius, noise, n_rays, grid, pthr, nthr, seed = cases.NMS2D_CASES[name]
    prob, dist = cases.create_random_data_2d(shape, radius, noise, n_rays, seed)
    fake = types.SimpleNamespace(config=types.SimpleNamespace(grid=(1, 1)), thresholds=thr)
    _, res = m2.StarDist2D._instances_from_prediction(fake, shape, prob, dist, return_labels=False)
    for k, v in res.items(): out["2d/dense/" + k] = v
    fake2 = types.SimpleNamespace(config=types.SimpleNamespace(grid=(2, 2)), thresholds=thr)
    rng = np.random.default_rng(0)
    pc = rng.uniform(size=prob[::2, ::2].shape + (4,)).astype(np.float32)
    _, res = m2.StarDist2D._instances_from_prediction(fake2, shape, prob[::2, ::2], dist[::2, ::2], prob_class=pc, return_labels=False, scale=dict(X=.5, Y=2.))
    for k, v in res.items(): out["2d/grid_scale_class/" + k] = v
    out["2d/grid_scale_class/prob_class_in"] = pc
    mask = prob > 0.92
    pts = np.stack(np.where(mask), 1)
    pcs = rng.uniform(size=(len(pts), 3)).astype(np.float32)
    _, res = m2.StarDist2D._instances_from_prediction(fake, shape, prob[mask], dist[mask], points=pts, prob_class=pcs, nms_thresh=0.4, return_labels=False)
    for k, v in res.items(): out["2d/sparse_class/" + k] = v
    out["2d/sparse_class/prob_class_in"] = pcs
    # ---- 3-D: dense with labels, sparse with scale
    name = "r32_noise01_thr01"
    shape, noise, n_rays, pthr, nthr, seed, aniso = cases.NMS3D_CASES[name]
    prob, dist = cases.create_random_data_3d(shape, noise, n_rays, seed)
    rays = rrays.Rays_GoldenSpiral(n_rays)
    fake = types.SimpleNamespace(config=types.SimpleNamespace(grid=(1, 1, 1), rays_json=rays.to_json()), thresholds=types.SimpleNamespace(prob=pthr, nms=nthr))
    labels, res = m3.StarDist3D._instances_from_prediction(fake, shape, prob, dist)
    out["3d/dense/labels"] = labels
    for k in ("dist", "points", "prob"): out["3d/dense/" + k] = res[k]
    mask = prob > pthr
    mask[:2] = mask[-2:] = False; mask[:, :2] = mask[:, -2:] = False; mask[:, :, :2] = mask[:, :, -2:] = False
    pts = np.stack(np.where(mask), 1)
    labels, res = m3.StarDist3D._instances_from_prediction(fake, shape, prob[mask], dist[mask], points=pts, nms_thresh=0.2, scale=dict(Z=1., Y=.5, X=2.))
    out["3d/sparse_scale/labels"] = labels
    for k in ("dist", "points", "prob", "rays_vertices"): out["3d/sparse_scale/" + k] = res[k]
    np.savez_compressed(os.path.join(HERE, "models_host.npz"), **out)
    print(len(out), "arrays;", {k: (v.shape, str(v.dtype)) for k, v in out.items() if 'labels' in k or 'coord' in k})
