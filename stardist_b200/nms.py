"""Non-maximum suppression front-end (numpy in / numpy out), device backed.

Public names, argument order and return values are those of stardist/nms.py -- _ind_prob_thresh (:6-17),
non_maximum_suppression (:77-132), non_maximum_suppression_sparse (:135-183), non_maximum_suppression_inds (:186-227) and the
3-D variants (:233-384) -- so the reference's call sites keep working.  The 2-D and 3-D front-ends share one implementation
here (`_dense`, `_sparse`); the only dimension specific step is the call into the device library.

Score order: the reference uses np.argsort(prob)[::-1], whose tie order is unspecified (unstable sort); here it is *defined* as
np.argsort(prob, kind='stable')[::-1] (the same definition as the device sort in csrc/candidates.cu and as the oracle).

The host logic is pinned on the CPU against the reference's own module (tests/test_cpu_oracle.py::
test_product_nms_front_end_equals_reference_modules, with the C entry points swapped for the reference's extensions).
"""
import numpy as np
from time import time
from .utils import _normalize_grid


def _ind_prob_thresh(prob, prob_thresh, b=2):
    """candidate mask: prob above the threshold and outside a border of b pixels (b: scalar, ((lo, hi), ...) per axis, or None)"""
    above = prob > prob_thresh
    if b is None:
        return above
    margins = ((b, b),) * prob.ndim if np.isscalar(b) else b
    interior = np.zeros_like(above)
    interior[tuple(slice(lo if lo > 0 else None, -hi if hi > 0 else None) for lo, hi in margins)] = True
    above &= interior
    return above


def _argsort_desc(scores):
    return np.argsort(scores, kind='stable')[::-1]


def _f32(x):
    return np.ascontiguousarray(x.astype(np.float32, copy=False))


def _say(verbose, msg):
    if verbose:
        print(msg, flush=True)


# ------------------------------------------------------------------------------------------ device entry points
def non_maximum_suppression_inds(dist, points, scores, thresh=0.5, use_bbox=True, use_kdtree=True, verbose=1):
    """Survivor mask of ray-convex polygons ALREADY sorted by descending score: P1 suppresses a later P2 if
    area(P1 ∩ P2) / min(area(P1), area(P2)) > thresh.  dist (n_poly, n_rays), points (n_poly, 2) -> bool (n_poly,)"""
    from .lib.stardist2d import c_non_max_suppression_inds
    assert dist.ndim == 2 and points.ndim == 2 and points.shape[0] == dist.shape[0]
    assert scores is None or len(scores) == dist.shape[0]
    # note the flag order of the 2-D entry point: kd-tree first, then bbox (stardist2d.cpp:396)
    return c_non_max_suppression_inds(_f32(dist), _f32(points), int(use_kdtree), int(use_bbox), int(verbose), np.float32(thresh))


def non_maximum_suppression_3d_inds(dist, points, rays, scores, thresh=0.5, use_bbox=True, use_kdtree=True, verbose=1):
    """Survivor mask of ray-convex polyhedra, in INPUT order: candidates are sorted by score here (nms.py:359-363), the
    device result is scattered back"""
    from .lib.stardist3d import c_non_max_suppression_inds
    assert dist.ndim == 2 and points.ndim == 2 and dist.shape[1] == len(rays) and points.shape[0] == dist.shape[0]
    n_poly = dist.shape[0]
    if scores is None:
        scores = np.ones(n_poly)
    assert len(scores) == n_poly
    order = _argsort_desc(scores)
    t0 = time()
    kept = c_non_max_suppression_inds(_f32(dist[order]), _f32(points[order]), _f32(rays.vertices),
                                      np.ascontiguousarray(rays.faces.astype(np.int32, copy=False)), _f32(scores[order]),
                                      int(use_bbox), int(use_kdtree), int(verbose), np.float32(thresh))
    _say(verbose, "NMS took %.4f s" % (time() - t0))
    survivors = np.ones(n_poly, bool)
    survivors[order] = kept
    return survivors


def _suppress(ndim, dist, points, scores, rays, **kw):
    if ndim == 2:
        return non_maximum_suppression_inds(dist, points, scores=scores, **kw)
    return non_maximum_suppression_3d_inds(dist, points, rays=rays, scores=scores, **kw)


# ------------------------------------------------------------------------------------------ shared front-ends
def _dense(ndim, dist, prob, rays, grid, b, nms_thresh, prob_thresh, use_bbox, use_kdtree, verbose):
    """prob [*spatial], dist [*spatial, n_rays] -> (points, prob, dist) of the survivors, points in image pixels (index * grid)"""
    dist, prob = np.asarray(dist), np.asarray(prob)
    assert prob.ndim == ndim and dist.ndim == ndim + 1 and prob.shape == dist.shape[:ndim]
    assert rays is None or dist.shape[-1] == len(rays)
    grid = _normalize_grid(grid, ndim)
    _say(verbose, "predicting instances with prob_thresh = %s and nms_thresh = %s" % (prob_thresh, nms_thresh))
    mask = _ind_prob_thresh(prob, prob_thresh, b)
    order = _argsort_desc(prob[mask])
    scores, rows = prob[mask][order], dist[mask][order]
    points = np.stack(np.where(mask), axis=1)[order] * np.array(grid).reshape((1, ndim))
    _say(verbose, "found %s candidates" % len(points))
    t0 = time()
    # the 2-D entry point has always been fed int32 pixel coordinates (nms.py:119)
    keep = _suppress(ndim, rows, points.astype(np.int32, copy=False) if ndim == 2 else points, scores, rays,
                     thresh=nms_thresh, use_bbox=use_bbox, use_kdtree=use_kdtree, verbose=verbose)
    _say(verbose, "keeping %s/%s candidates, NMS took %.4f s" % (np.count_nonzero(keep), len(keep), time() - t0))
    return points[keep], scores[keep], rows[keep]


def _sparse(ndim, dist, prob, points, rays, nms_thresh, use_kdtree, verbose):
    """candidate lists in any order -> (points, prob, dist, original indices) of the survivors, by descending score"""
    dist, prob, points = np.asarray(dist), np.asarray(prob), np.asarray(points)
    assert dist.ndim == 2 and prob.ndim == 1 and points.ndim == 2 and points.shape[-1] == ndim
    assert len(prob) == len(dist) == len(points) and (rays is None or dist.shape[-1] == len(rays))
    _say(verbose, "predicting instances with nms_thresh = %s" % nms_thresh)
    order = _argsort_desc(prob)
    scores, rows, pts = prob[order], dist[order], points[order]
    t0 = time()
    keep = _suppress(ndim, rows, pts, scores, rays, thresh=nms_thresh, use_kdtree=use_kdtree, verbose=verbose)
    _say(verbose, "keeping %s/%s candidates, NMS took %.4f s" % (np.count_nonzero(keep), len(keep), time() - t0))
    return pts[keep], scores[keep], rows[keep], order[keep]


# ------------------------------------------------------------------------------------------ public API (reference names)
def non_maximum_suppression(dist, prob, grid=(1, 1), b=2, nms_thresh=0.5, prob_thresh=0.5,
                            use_bbox=True, use_kdtree=True, verbose=False):
    """2-D, dense maps: dist (Ny, Nx, n_rays), prob (Ny, Nx) -> points, prob, dist of the retained polygons"""
    return _dense(2, dist, prob, None, grid, b, nms_thresh, prob_thresh, use_bbox, use_kdtree, verbose)


def non_maximum_suppression_sparse(dist, prob, points, b=2, nms_thresh=0.5, use_bbox=True, use_kdtree=True, verbose=False):
    """2-D, candidate lists: dist (n, n_rays), prob (n,), points (n, 2) -> points, prob, dist, inds with points == points_in[inds]"""
    return _sparse(2, dist, prob, points, None, nms_thresh, use_kdtree, verbose)


def non_maximum_suppression_3d(dist, prob, rays, grid=(1, 1, 1), b=2, nms_thresh=0.5, prob_thresh=0.5,
                               use_bbox=True, use_kdtree=True, verbose=False):
    """3-D, dense maps: dist (Nz, Ny, Nx, n_rays), prob (Nz, Ny, Nx) -> points, prob, dist of the retained polyhedra"""
    return _dense(3, dist, prob, rays, grid, b, nms_thresh, prob_thresh, use_bbox, use_kdtree, verbose)


def non_maximum_suppression_3d_sparse(dist, prob, points, rays, b=2, nms_thresh=0.5, use_kdtree=True, verbose=False):
    """3-D, candidate lists -> points, prob, dist, inds of the retained polyhedra"""
    return _sparse(3, dist, prob, points, rays, nms_thresh, use_kdtree, verbose)
