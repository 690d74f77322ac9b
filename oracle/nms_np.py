"""numpy restatement of stardist/nms.py on top of the reference's compiled C++ (oracle/_ref).
TEST INFRASTRUCTURE ONLY.

Follows: _ind_prob_thresh nms.py:6-17; non_maximum_suppression :77-132;
non_maximum_suppression_sparse :135-183; non_maximum_suppression_inds :186-227;
3D variants :233-384.  Tie order: np.argsort(..., kind='stable')[::-1] (the reference's default
argsort is unstable; see DESIGN.md "score order").
"""
import numpy as np
from . import ref_ext


def _ind_prob_thresh(prob, prob_thresh, b=2):
    if b is not None and np.isscalar(b):
        b = ((b, b),) * prob.ndim
    ind_thresh = prob > prob_thresh
    if b is not None:
        _ind_thresh = np.zeros_like(ind_thresh)
        ss = tuple(slice(_bs[0] if _bs[0] > 0 else None, -_bs[1] if _bs[1] > 0 else None) for _bs in b)
        _ind_thresh[ss] = True
        ind_thresh &= _ind_thresh
    return ind_thresh


def argsort_desc(x):
    return np.argsort(x, kind='stable')[::-1]


def _prep(x, dtype):
    return np.ascontiguousarray(x.astype(dtype, copy=False))


def non_maximum_suppression_inds(dist, points, scores=None, thresh=0.5, use_bbox=True, use_kdtree=True, verbose=0):
    return ref_ext.stardist2d().c_non_max_suppression_inds(_prep(dist, np.float32), _prep(points, np.float32),
                                                           int(use_kdtree), int(use_bbox), int(verbose), np.float32(thresh))


def non_maximum_suppression(dist, prob, grid=(1, 1), b=2, nms_thresh=0.5, prob_thresh=0.5, use_bbox=True, use_kdtree=True):
    mask = _ind_prob_thresh(prob, prob_thresh, b)
    points = np.stack(np.where(mask), axis=1)
    dist = dist[mask]; scores = prob[mask]
    ind = argsort_desc(scores)
    dist, scores, points = dist[ind], scores[ind], points[ind]
    points = points * np.array(grid).reshape((1, 2))
    inds = non_maximum_suppression_inds(dist, points.astype(np.int32, copy=False), scores, thresh=nms_thresh,
                                        use_bbox=use_bbox, use_kdtree=use_kdtree)
    return points[inds], scores[inds], dist[inds]


def non_maximum_suppression_sparse(dist, prob, points, nms_thresh=0.5, use_bbox=True, use_kdtree=True):
    inds_original = np.arange(len(prob))
    _sorted = argsort_desc(prob)
    probi, disti, pointsi = prob[_sorted], dist[_sorted], points[_sorted]
    inds_original = inds_original[_sorted]
    inds = non_maximum_suppression_inds(disti, pointsi, scores=probi, thresh=nms_thresh, use_kdtree=use_kdtree)
    return pointsi[inds], probi[inds], disti[inds], inds_original[inds]


# ---- 3D (nms.py:285-384) -----------------------------------------------------------------
def non_maximum_suppression_3d_inds(dist, points, rays, scores, thresh=0.5, use_bbox=True, use_kdtree=True, verbose=0):
    # nms.py:359-363 sorts again by scores (descending, here stable) and maps the result back
    ind = argsort_desc(scores)
    survivors = np.zeros(len(ind), bool)
    d, p, s = dist[ind], points[ind], scores[ind]
    inds = ref_ext.stardist3d().c_non_max_suppression_inds(_prep(d, np.float32), _prep(p, np.float32),
                                                           _prep(rays.vertices, np.float32), _prep(rays.faces, np.int32),
                                                           _prep(s, np.float32), int(use_bbox), int(use_kdtree), int(verbose),
                                                           np.float32(thresh))
    survivors[ind] = inds
    return survivors


def non_maximum_suppression_3d_sparse(dist, prob, points, rays, nms_thresh=0.5, use_bbox=True, use_kdtree=True):
    inds_original = np.arange(len(prob))
    _sorted = argsort_desc(prob)
    probi, disti, pointsi = prob[_sorted], dist[_sorted], points[_sorted]
    inds_original = inds_original[_sorted]
    inds = non_maximum_suppression_3d_inds(disti, pointsi, rays, probi, thresh=nms_thresh, use_bbox=use_bbox, use_kdtree=use_kdtree)
    return pointsi[inds], probi[inds], disti[inds], inds_original[inds]
