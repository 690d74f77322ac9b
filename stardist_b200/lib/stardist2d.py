"""Drop-in for `stardist.lib.stardist2d` (reference: stardist/lib/stardist2d.cpp:621-632).

c_non_max_suppression_inds(dist f32[n,R] C-contig, points f32[n,2], use_kdtree:int, use_bbox:int,
                           verbose:int, thresh:f32) -> bool[n]         ("O!O!iiif", stardist2d.cpp:396)
Inputs must be sorted by descending score (as stardist/nms.py:186-227 guarantees).
"""
import numpy as np
from .. import _lib as L


def c_non_max_suppression_inds(dist, points, use_kdtree, use_bbox, verbose, thresh):
    if not (isinstance(dist, np.ndarray) and isinstance(points, np.ndarray)):
        raise TypeError("dist and points must be numpy arrays")
    if dist.dtype != np.float32 or points.dtype != np.float32:
        raise TypeError("dist and points must be float32")
    if dist.ndim != 2 or points.ndim != 2 or points.shape[1] != 2 or points.shape[0] != dist.shape[0]:
        raise ValueError("expected dist (n,R) and points (n,2)")
    lib = L.require_cuda()
    dist = np.ascontiguousarray(dist); points = np.ascontiguousarray(points)
    n, R = dist.shape
    result = np.zeros(n, dtype=np.bool_)
    if n > 0:
        L.check(lib._LIB_non_maximum_suppression_2d(L.ptr(dist), L.ptr(points), n, R, float(thresh),
                                                   int(use_bbox), int(use_kdtree), int(verbose), L.ptr(result)))
    return result
