"""CPU restatement of StarDist2D.predict_instances (sparse path) -- TEST INFRASTRUCTURE ONLY.

Follows stardist/models/base.py:371-443 (_predict_setup: axes, reflect-pad at the end to a multiple
of 2^depth*grid), :541-633 (_predict_sparse_generator: prob = ch0, dist = max(1e-3, dist),
_ind_prob_thresh with b=2 on the padded map, points = idx*grid, filter_points), and
stardist/models/model2d.py:512-563 (_instances_from_prediction: NMS -> polygons_to_label ->
dist_to_coord).  Network = oracle.unet_torch (torch-CPU fp32 stand-in for TF).
"""
import numpy as np
from . import unet_torch, nms_np, geom2d_np


def _pad(x, div_by):
    pads = [(0, (d - s % d) % d) for s, d in zip(x.shape, div_by)]
    return np.pad(x, pads, mode='reflect'), pads


def predict(config, weights, img):
    """img [H,W] (single channel) -> padded prob [Hp/g, Wp/g], dist [.., R], orig shape"""
    g = config.grid
    div = tuple(p ** config.unet_n_depth * gg for p, gg in zip(config.unet_pool, g))
    x, pads = _pad(np.asarray(img, np.float32), div)
    prob, dist = unet_torch.forward(config, weights, x[np.newaxis, ..., np.newaxis])
    return prob[0], dist[0], pads


def candidates(config, prob, dist, pads, img_shape, prob_thresh, b=2):
    dist = np.maximum(np.float32(1e-3), dist)
    inds = nms_np._ind_prob_thresh(prob, prob_thresh, b=b)
    proba = prob[inds].copy(); dista = dist[inds].copy()
    points = np.stack(np.where(inds), axis=1) * np.array(config.grid).reshape(1, 2)
    bounds = np.array(img_shape)            # padded_shape - pad
    idx = np.where(np.all(points < bounds, 1))
    return proba[idx], dista[idx], points[idx]


def instances(config, img_shape, proba, dista, points, nms_thresh):
    points, probi, disti, indsi = nms_np.non_maximum_suppression_sparse(dista, proba, points, nms_thresh=nms_thresh)
    labels = geom2d_np.polygons_to_label(disti, points, prob=probi, shape=img_shape)
    coord = geom2d_np.dist_to_coord(disti, points)
    return labels, dict(coord=coord, points=points, prob=probi, dist=disti)


def predict_instances(config, weights, img, prob_thresh, nms_thresh, cand_from=None):
    """cand_from: a product model -- its *device* prob/dist maps are used as the network output so
    that the integer post-processing is compared bit-exactly on identical floats (the network
    itself is compared separately with a tolerance)."""
    if cand_from is not None:
        prob, dist = cand_from._last_maps()
        pads = None
    else:
        prob, dist, pads = predict(config, weights, img)
    proba, dista, points = candidates(config, prob, dist, pads, img.shape, prob_thresh)
    return instances(config, img.shape, proba, dista, points, nms_thresh)


def quantile_prob_thresh(model, img, q):
    """prob threshold giving a fixed candidate fraction for a random-init network"""
    prob, dist = model.predict(img)
    return float(np.quantile(prob, q))
