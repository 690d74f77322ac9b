"""3D geometry of the prediction path, device backed.

Mirrors stardist/geometry/geom3d.py: polyhedron_to_label (:100-198), dist_to_coord3D (:261-275).
"""
import numpy as np


RENDER_MODES = ("full", "kernel", "hull", "bbox", "debug")      # index = render_mode of the C entry point


def _as_rows(a, what):
    """one polyhedron may be given as a flat vector"""
    a = np.asanyarray(a)
    return a.reshape(1, -1) if a.ndim == 1 else a


def polyhedron_to_label(dist, points, rays, shape, prob=None, thr=-np.inf, labels=None, mode="full", verbose=True, overlap_label=None):
    """Label image of shape `shape` from star-convex polyhedra (API of stardist/geometry/geom3d.py:100-198).

    dist (n, n_rays) ray lengths, points (n, 3) centres, rays a Rays object; prob (n,) scores (default all one) select the
    polyhedra with prob >= thr and fix the painting order -- descending score, the first polyhedron covering a voxel wins;
    labels (n,) ids to write (default 1..n); mode one of RENDER_MODES; overlap_label: value for voxels covered more than once.
    Rendering: csrc/label3d.cu through the reference-signature entry point."""
    from ..lib.stardist3d import c_polyhedron_to_label
    n = len(points)
    if n == 0:
        verbose and print("warning: empty list of points (returning background-only image)")
        return np.zeros(shape, np.uint16)
    dist, points = _as_rows(dist, "dist"), _as_rows(points, "points")
    labels = np.arange(1, n + 1) if labels is None else np.asanyarray(labels)
    prob = np.ones(n) if prob is None else np.asanyarray(prob)
    # argument checks, with the reference's messages
    problems = (
        (np.amin(dist) <= 0, ValueError("distance array should be positive!")),
        (dist.ndim != 2, ValueError("dist should be 2 dimensional but has shape %s" % str(dist.shape))),
        (dist.ndim == 2 and dist.shape[1] != len(rays), ValueError("inconsistent number of rays!")),
        (len(prob) != n, ValueError("len(prob) != len(points)")),
        (len(labels) != n, ValueError("len(labels) != len(points)")),
        (mode not in RENDER_MODES, KeyError("Unknown render mode '%s' , allowed:  %s" % (mode, RENDER_MODES))),
    )
    for failed, err in problems:
        if failed:
            raise err
    # selection and painting order in one index vector: the survivors of the threshold, best score first
    # (the reference's argsort()[::-1] is unstable on ties; defined here as the stable sort)
    chosen = np.flatnonzero(prob >= thr)
    if chosen.size == 0:
        verbose and print("warning: no points found with probability>= {thr:.4f} (returning background-only image)".format(thr=thr))
        return np.zeros(shape, np.uint16)
    chosen = chosen[np.argsort(prob[chosen], kind='stable')[::-1]]
    return c_polyhedron_to_label(np.ascontiguousarray(dist[chosen], np.float32), np.ascontiguousarray(points[chosen], np.float32),
                                 np.ascontiguousarray(rays.vertices, np.float32), np.ascontiguousarray(rays.faces, np.int32),
                                 np.ascontiguousarray(labels[chosen], np.int32), np.int32(RENDER_MODES.index(mode)), np.int32(verbose),
                                 np.int32(overlap_label is not None), np.int32(0 if overlap_label is None else overlap_label), shape)


def dist_to_coord3D(dist, points, rays_vertices):
    """ converts dist/points/rays_vertices to list of coords """
    dist = np.asarray(dist)
    points = np.asarray(points)
    rays_vertices = np.asarray(rays_vertices)
    if not all((len(dist) == len(points), dist.ndim == 2, points.ndim == 2,
                points.shape[-1] == 3, rays_vertices.shape[-1] == 3, dist.shape[-1] == len(rays_vertices))):
        raise ValueError(f"Wrong shapes! dist -> (m,n) points -> (m,3) rays_vertices -> (m,)")
    return points[:, np.newaxis] + dist[..., np.newaxis] * rays_vertices
