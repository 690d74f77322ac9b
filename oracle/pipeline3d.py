"""CPU restatement of StarDist3D.predict_instances (sparse path) -- TEST INFRASTRUCTURE ONLY.

Follows stardist/models/base.py:371-443,541-633 (as pipeline2d) and stardist/models/model3d.py:589-674
(_instances_from_prediction: non_maximum_suppression_3d_sparse -> polyhedron_to_label (reference C++,
oracle/_ref) -> relabel_sequential).  Network = oracle.unet_torch (torch-CPU fp32 stand-in for TF).
"""
import numpy as np
from . import unet_torch, nms_np, ref_ext


def relabel_sequential(label_field, offset=1):
    # stardist/matching.py:319-406
    labels = np.unique(label_field); labels0 = labels[labels != 0]
    new_max = offset - 1 + len(labels0)
    out_t = label_field.dtype
    req = np.min_scalar_type(new_max)
    if np.dtype(req).itemsize > np.dtype(out_t).itemsize: out_t = req
    fwd = np.zeros(int(label_field.max()) + 1, dtype=out_t)
    fwd[labels0] = np.arange(offset, new_max + 1)
    return fwd[label_field]


def polyhedron_to_label(dist, points, rays, shape, prob):
    # stardist/geometry/geom3d.py:100-198 (mode "full", no overlap label), stable descending order
    if len(points) == 0:
        return np.zeros(shape, np.uint16)
    ind = np.argsort(prob, kind='stable')[::-1]
    d, p = dist[ind], points[ind]
    labels = np.arange(1, len(points) + 1)[ind]
    prep = lambda x, t: np.ascontiguousarray(np.asarray(x).astype(t, copy=False))
    return ref_ext.stardist3d().c_polyhedron_to_label(prep(d, np.float32), prep(p, np.float32), prep(rays.vertices, np.float32),
                                                     prep(rays.faces, np.int32), prep(labels, np.int32), np.int32(0), np.int32(0),
                                                     np.int32(0), np.int32(0), tuple(int(s) for s in shape))


def candidates(config, prob, dist, img_shape, prob_thresh, b=2):
    dist = np.maximum(np.float32(1e-3), dist)
    inds = nms_np._ind_prob_thresh(prob, prob_thresh, b=b)
    proba = prob[inds].copy(); dista = dist[inds].copy()
    points = np.stack(np.where(inds), axis=1) * np.array(config.grid).reshape(1, 3)
    idx = np.where(np.all(points < np.array(img_shape), 1))
    return proba[idx], dista[idx], points[idx]


def instances(config, rays, img_shape, proba, dista, points, nms_thresh):
    points, probi, disti, indsi = nms_np.non_maximum_suppression_3d_sparse(dista, proba, points, rays, nms_thresh=nms_thresh)
    labels = polyhedron_to_label(disti, points, rays, img_shape, probi)
    labels = relabel_sequential(labels)
    return labels, dict(dist=disti, points=points, prob=probi)


def predict(config, weights, img):
    """img [D,H,W] (single channel) -> prob / dist of the reflect-padded volume (base.py:371-443; the volume is padded
    at the end to a multiple of pool^depth*grid for the U-Net, of grid for the ResNet, model3d.py:676-688)"""
    if getattr(config, 'backbone', 'unet') == 'resnet':
        div = tuple(config.grid)
    else:
        div = tuple(p ** config.unet_n_depth * g for p, g in zip(config.unet_pool, config.grid))
    x = np.asarray(img, np.float32)
    x = np.pad(x, [(0, (d - s % d) % d) for s, d in zip(x.shape, div)], mode='reflect')
    prob, dist = unet_torch.forward(config, weights, x[np.newaxis, ..., np.newaxis])
    return prob[0], dist[0]


def predict_instances(config, rays, img, prob_thresh, nms_thresh, cand_from=None, weights=None):
    """cand_from: a product model whose device prob/dist maps are used (integer post-processing compared on identical
    floats); otherwise the torch-CPU network runs on `weights`"""
    if cand_from is not None:
        prob, dist = cand_from._last_maps()
    else:
        prob, dist = predict(config, weights, img)
    proba, dista, points = candidates(config, prob, dist, img.shape, prob_thresh)
    return instances(config, rays, img.shape, proba, dista, points, nms_thresh)
