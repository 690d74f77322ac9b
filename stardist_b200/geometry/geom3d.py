"""3D geometry of the prediction path, device backed.

Mirrors stardist/geometry/geom3d.py: polyhedron_to_label (:100-198), dist_to_coord3D (:261-275).
"""
import numpy as np


def polyhedron_to_label(dist, points, rays, shape, prob=None, thr=-np.inf, labels=None, mode="full", verbose=True, overlap_label=None):
    """creates labeled image from stardist representations (see the reference docstring)

    mode: "full", "kernel", "hull", "bbox" or "debug"
    """
    from ..lib.stardist3d import c_polyhedron_to_label
    if len(points) == 0:
        if verbose:
            print("warning: empty list of points (returning background-only image)")
        return np.zeros(shape, np.uint16)
    dist = np.asanyarray(dist)
    points = np.asanyarray(points)
    if dist.ndim == 1:
        dist = dist.reshape(1, -1)
    if points.ndim == 1:
        points = points.reshape(1, -1)
    if labels is None:
        labels = np.arange(1, len(points) + 1)
    labels = np.asanyarray(labels)
    if np.amin(dist) <= 0:
        raise ValueError("distance array should be positive!")
    prob = np.ones(len(points)) if prob is None else np.asanyarray(prob)
    if dist.ndim != 2:
        raise ValueError("dist should be 2 dimensional but has shape %s" % str(dist.shape))
    if dist.shape[1] != len(rays):
        raise ValueError("inconsistent number of rays!")
    if len(prob) != len(points):
        raise ValueError("len(prob) != len(points)")
    if len(labels) != len(points):
        raise ValueError("len(labels) != len(points)")
    modes = {"full": 0, "kernel": 1, "hull": 2, "bbox": 3, "debug": 4}
    if mode not in modes:
        raise KeyError("Unknown render mode '%s' , allowed:  %s" % (mode, tuple(modes.keys())))
    lbl = np.zeros(shape, np.uint16)
    # filter points
    ind = np.where(prob >= thr)[0]
    if len(ind) == 0:
        if verbose:
            print("warning: no points found with probability>= {thr:.4f} (returning background-only image)".format(thr=thr))
        return lbl
    prob = prob[ind]
    points = points[ind]
    dist = dist[ind]
    labels = labels[ind]
    # sort points with decreasing probability (stable definition of the reference's argsort()[::-1])
    ind = np.argsort(prob, kind='stable')[::-1]
    points = points[ind]
    dist = dist[ind]
    labels = labels[ind]

    def _prep(x, dtype):
        return np.ascontiguousarray(x.astype(dtype, copy=False))
    return c_polyhedron_to_label(_prep(dist, np.float32), _prep(points, np.float32), _prep(rays.vertices, np.float32),
                                 _prep(rays.faces, np.int32), _prep(labels, np.int32), np.int32(modes[mode]),
                                 np.int32(verbose), np.int32(overlap_label is not None),
                                 np.int32(0 if overlap_label is None else overlap_label), shape)


def dist_to_coord3D(dist, points, rays_vertices):
    """ converts dist/points/rays_vertices to list of coords """
    dist = np.asarray(dist)
    points = np.asarray(points)
    rays_vertices = np.asarray(rays_vertices)
    if not all((len(dist) == len(points), dist.ndim == 2, points.ndim == 2,
                points.shape[-1] == 3, rays_vertices.shape[-1] == 3, dist.shape[-1] == len(rays_vertices))):
        raise ValueError(f"Wrong shapes! dist -> (m,n) points -> (m,3) rays_vertices -> (m,)")
    return points[:, np.newaxis] + dist[..., np.newaxis] * rays_vertices
