"""Non-maximum suppression front-end (numpy in / numpy out), device backed.

Mirrors stardist/nms.py: _ind_prob_thresh (:6-17), non_maximum_suppression (:77-132),
non_maximum_suppression_sparse (:135-183), non_maximum_suppression_inds (:186-227) and the 3D
variants (:233-384).  Score order: the reference uses np.argsort(prob)[::-1], whose tie order
is unspecified (unstable sort); here it is *defined* as np.argsort(prob, kind='stable')[::-1]
(the same definition as the device sort in csrc/candidates.cu and as the oracle).
"""
import numpy as np
from time import time
from .utils import _normalize_grid


def _ind_prob_thresh(prob, prob_thresh, b=2):
    if b is not None and np.isscalar(b):
        b = ((b, b),) * prob.ndim
    ind_thresh = prob > prob_thresh
    if b is not None:
        _ind_thresh = np.zeros_like(ind_thresh)
        ss = tuple(slice(_bs[0] if _bs[0] > 0 else None,
                         -_bs[1] if _bs[1] > 0 else None) for _bs in b)
        _ind_thresh[ss] = True
        ind_thresh &= _ind_thresh
    return ind_thresh


def _argsort_desc(scores):
    return np.argsort(scores, kind='stable')[::-1]


def non_maximum_suppression(dist, prob, grid=(1, 1), b=2, nms_thresh=0.5, prob_thresh=0.5,
                            use_bbox=True, use_kdtree=True, verbose=False):
    """Non-Maximum-Supression of 2D polygons

    Retains only polygons whose overlap is smaller than nms_thresh

    dist.shape = (Ny,Nx, n_rays)
    prob.shape = (Ny,Nx)

    returns the retained points, probabilities, and distances:

    points, prob, dist = non_maximum_suppression(dist, prob, ....
    """
    assert prob.ndim == 2 and dist.ndim == 3 and prob.shape == dist.shape[:2]
    dist = np.asarray(dist)
    prob = np.asarray(prob)
    grid = _normalize_grid(grid, 2)
    mask = _ind_prob_thresh(prob, prob_thresh, b)
    points = np.stack(np.where(mask), axis=1)
    dist = dist[mask]
    scores = prob[mask]
    ind = _argsort_desc(scores)
    dist = dist[ind]
    scores = scores[ind]
    points = points[ind]
    points = (points * np.array(grid).reshape((1, 2)))
    if verbose:
        t = time()
    inds = non_maximum_suppression_inds(dist, points.astype(np.int32, copy=False), scores=scores,
                                        use_bbox=use_bbox, use_kdtree=use_kdtree,
                                        thresh=nms_thresh, verbose=verbose)
    if verbose:
        print("keeping %s/%s polygons" % (np.count_nonzero(inds), len(inds)))
        print("NMS took %.4f s" % (time() - t))
    return points[inds], scores[inds], dist[inds]


def non_maximum_suppression_sparse(dist, prob, points, b=2, nms_thresh=0.5,
                                   use_bbox=True, use_kdtree=True, verbose=False):
    """Non-Maximum-Supression of 2D polygons from a list of dists, probs (scores), and points

    dist.shape = (n_polys, n_rays), prob.shape = (n_polys,), points.shape = (n_polys,2)
    returns the retained instances (pointsi, probi, disti, indsi) with pointsi = points[indsi] ...
    """
    dist = np.asarray(dist)
    prob = np.asarray(prob)
    points = np.asarray(points)
    assert dist.ndim == 2 and prob.ndim == 1 and points.ndim == 2 and \
        points.shape[-1] == 2 and len(prob) == len(dist) == len(points)
    verbose and print("predicting instances with nms_thresh = {nms_thresh}".format(nms_thresh=nms_thresh), flush=True)
    inds_original = np.arange(len(prob))
    _sorted = _argsort_desc(prob)
    probi = prob[_sorted]
    disti = dist[_sorted]
    pointsi = points[_sorted]
    inds_original = inds_original[_sorted]
    if verbose:
        print("non-maximum suppression...")
        t = time()
    inds = non_maximum_suppression_inds(disti, pointsi, scores=probi, thresh=nms_thresh, use_kdtree=use_kdtree, verbose=verbose)
    if verbose:
        print("keeping %s/%s polyhedra" % (np.count_nonzero(inds), len(inds)))
        print("NMS took %.4f s" % (time() - t))
    return pointsi[inds], probi[inds], disti[inds], inds_original[inds]


def non_maximum_suppression_inds(dist, points, scores, thresh=0.5, use_bbox=True, use_kdtree=True, verbose=1):
    """
    Applies non maximum supression to ray-convex polygons given by dists and points
    sorted by scores and IoU threshold

    P1 will suppress P2, if IoU(P1,P2) > thresh
    with IoU(P1,P2) = Ainter(P1,P2) / min(A(P1),A(P2))

    dist.shape = (n_poly, n_rays), point.shape = (n_poly, 2), score.shape = (n_poly,)
    returns indices of selected polygons
    """
    from .lib.stardist2d import c_non_max_suppression_inds
    assert dist.ndim == 2
    assert points.ndim == 2
    n_poly = dist.shape[0]
    if scores is None:
        scores = np.ones(n_poly)
    assert len(scores) == n_poly
    assert points.shape[0] == n_poly

    def _prep(x, dtype):
        return np.ascontiguousarray(x.astype(dtype, copy=False))

    inds = c_non_max_suppression_inds(_prep(dist, np.float32),
                                      _prep(points, np.float32),
                                      int(use_kdtree),
                                      int(use_bbox),
                                      int(verbose),
                                      np.float32(thresh))
    return inds


#########  3D  (stardist/nms.py:233-384)

def non_maximum_suppression_3d(dist, prob, rays, grid=(1, 1, 1), b=2, nms_thresh=0.5, prob_thresh=0.5, use_bbox=True, use_kdtree=True, verbose=False):
    """Non-Maximum-Supression of 3D polyhedra

    dist.shape = (Nz,Ny,Nx, n_rays), prob.shape = (Nz,Ny,Nx)
    returns the retained points, probabilities, and distances
    """
    dist = np.asarray(dist)
    prob = np.asarray(prob)
    assert prob.ndim == 3 and dist.ndim == 4 and dist.shape[-1] == len(rays) and prob.shape == dist.shape[:3]
    grid = _normalize_grid(grid, 3)
    verbose and print("predicting instances with prob_thresh = {prob_thresh} and nms_thresh = {nms_thresh}".format(prob_thresh=prob_thresh, nms_thresh=nms_thresh), flush=True)
    ind_thresh = _ind_prob_thresh(prob, prob_thresh, b)
    points = np.stack(np.where(ind_thresh), axis=1)
    verbose and print("found %s candidates" % len(points))
    probi = prob[ind_thresh]
    disti = dist[ind_thresh]
    _sorted = _argsort_desc(probi)
    probi = probi[_sorted]
    disti = disti[_sorted]
    points = points[_sorted]
    verbose and print("non-maximum suppression...")
    points = (points * np.array(grid).reshape((1, 3)))
    inds = non_maximum_suppression_3d_inds(disti, points, rays=rays, scores=probi, thresh=nms_thresh,
                                           use_bbox=use_bbox, use_kdtree=use_kdtree, verbose=verbose)
    verbose and print("keeping %s/%s polyhedra" % (np.count_nonzero(inds), len(inds)))
    return points[inds], probi[inds], disti[inds]


def non_maximum_suppression_3d_sparse(dist, prob, points, rays, b=2, nms_thresh=0.5, use_kdtree=True, verbose=False):
    """Non-Maximum-Supression of 3D polyhedra from a list of dists, probs and points

    returns the retained instances (pointsi, probi, disti, indsi) with pointsi = points[indsi] ...
    """
    dist = np.asarray(dist)
    prob = np.asarray(prob)
    points = np.asarray(points)
    assert dist.ndim == 2 and prob.ndim == 1 and points.ndim == 2 and \
        dist.shape[-1] == len(rays) and points.shape[-1] == 3 and len(prob) == len(dist) == len(points)
    verbose and print("predicting instances with nms_thresh = {nms_thresh}".format(nms_thresh=nms_thresh), flush=True)
    inds_original = np.arange(len(prob))
    _sorted = _argsort_desc(prob)
    probi = prob[_sorted]
    disti = dist[_sorted]
    pointsi = points[_sorted]
    inds_original = inds_original[_sorted]
    verbose and print("non-maximum suppression...")
    inds = non_maximum_suppression_3d_inds(disti, pointsi, rays=rays, scores=probi, thresh=nms_thresh, use_kdtree=use_kdtree, verbose=verbose)
    verbose and print("keeping %s/%s polyhedra" % (np.count_nonzero(inds), len(inds)))
    return pointsi[inds], probi[inds], disti[inds], inds_original[inds]


def non_maximum_suppression_3d_inds(dist, points, rays, scores, thresh=0.5, use_bbox=True, use_kdtree=True, verbose=1):
    """
    Applies non maximum supression to ray-convex polyhedra given by dists and rays
    sorted by scores and IoU threshold; returns the boolean survivor mask (in input order)
    """
    from .lib.stardist3d import c_non_max_suppression_inds
    assert dist.ndim == 2
    assert points.ndim == 2
    assert dist.shape[1] == len(rays)
    n_poly = dist.shape[0]
    if scores is None:
        scores = np.ones(n_poly)
    assert len(scores) == n_poly
    assert points.shape[0] == n_poly
    # sort scores descendingly (nms.py:359-363)
    ind = _argsort_desc(scores)
    survivors = np.ones(n_poly, bool)
    dist = dist[ind]
    points = points[ind]
    scores = scores[ind]

    def _prep(x, dtype):
        return np.ascontiguousarray(x.astype(dtype, copy=False))
    if verbose:
        t = time()
    survivors[ind] = c_non_max_suppression_inds(_prep(dist, np.float32), _prep(points, np.float32),
                                                _prep(rays.vertices, np.float32), _prep(rays.faces, np.int32),
                                                _prep(scores, np.float32), int(use_bbox), int(use_kdtree),
                                                int(verbose), np.float32(thresh))
    if verbose:
        print("NMS took %.4f s" % (time() - t))
    return survivors
