"""numpy restatement of stardist/geometry/geom2d.py (dist_to_coord :130-146,
polygons_to_label_coord :149-166, polygons_to_label :169-197) and of the third-party
skimage.draw.polygon it calls (un-vendored, setup.py:141; rule of skimage >= 0.18:
skimage/draw/_draw.pyx::_polygon + skimage/_shared/geometry.pyx::point_in_polygon).
TEST INFRASTRUCTURE ONLY.  [polygon rule: parity UNPINNED -- no skimage in this image]
"""
import numpy as np


def ray_angles(n_rays=32):
    return np.linspace(0, 2 * np.pi, n_rays, endpoint=False)


def dist_to_coord(dist, points, scale_dist=(1, 1)):
    dist = np.asarray(dist); points = np.asarray(points)
    n_rays = dist.shape[1]
    phis = ray_angles(n_rays)
    coord = (dist[:, np.newaxis] * np.array([np.sin(phis), np.cos(phis)])).astype(np.float32)
    coord *= np.asarray(scale_dist).reshape(1, 2, 1)
    coord += points[..., np.newaxis]
    return coord


def polygon(r, c, shape):
    """skimage.draw.polygon(r, c, shape) -> (rr, cc)"""
    r = np.asanyarray(r); c = np.asanyarray(c)
    minr = int(max(0, r.min())); maxr = int(np.ceil(r.max()))
    minc = int(max(0, c.min())); maxc = int(np.ceil(c.max()))
    if shape is not None:
        maxr = min(shape[0] - 1, maxr); maxc = min(shape[1] - 1, maxc)
    if maxr < minr or maxc < minc:
        return np.zeros(0, np.intp), np.zeros(0, np.intp)
    yp = np.ascontiguousarray(r, 'float64'); xp = np.ascontiguousarray(c, 'float64')
    rs = np.arange(minr, maxr + 1, dtype=np.float64); cs = np.arange(minc, maxc + 1, dtype=np.float64)
    Y, X = np.meshgrid(rs, cs, indexing='ij')
    eps = 1e-12
    l_cross = np.zeros(Y.shape, np.int64); r_cross = np.zeros(Y.shape, np.int64)
    vertex = np.zeros(Y.shape, bool)
    x1 = xp[-1] - X; y1 = yp[-1] - Y
    with np.errstate(divide='ignore', invalid='ignore'):
        for i in range(len(xp)):
            x0 = xp[i] - X; y0 = yp[i] - Y
            isv = (-eps < x0) & (x0 < eps) & (-eps < y0) & (y0 < eps)
            live = ~vertex          # the C loop returns at the first vertex hit
            vertex |= isv
            live &= ~isv
            q = (x0 * y1 - x1 * y0) / (y1 - y0)
            r_cross += (live & ((y0 > 0) != (y1 > 0)) & (q > 0))
            l_cross += (live & ((y0 < 0) != (y1 < 0)) & (q < 0))
            x1, y1 = x0, y0
    inside = vertex | ((r_cross & 1) != (l_cross & 1)) | ((r_cross & 1) == 1)
    rr, cc = np.nonzero(inside)
    return (rr + minr).astype(np.intp), (cc + minc).astype(np.intp)


def polygons_to_label_coord(coord, shape, labels=None):
    coord = np.asarray(coord)
    if labels is None: labels = np.arange(len(coord))
    lbl = np.zeros(shape, np.int32)
    for i, c in zip(labels, coord):
        rr, cc = polygon(*c, shape)
        lbl[rr, cc] = i + 1
    return lbl


def polygons_to_label(dist, points, shape, prob=None, thr=-np.inf, scale_dist=(1, 1)):
    dist = np.asarray(dist); points = np.asarray(points)
    prob = np.inf * np.ones(len(points)) if prob is None else np.asarray(prob)
    ind = prob > thr
    points, dist, prob = points[ind], dist[ind], prob[ind]
    ind = np.argsort(prob, kind='stable')
    points, dist = points[ind], dist[ind]
    coord = dist_to_coord(dist, points, scale_dist=scale_dist)
    return polygons_to_label_coord(coord, shape=shape, labels=ind)
