"""Synthetic ground truth helpers for bench / tests -- TEST INFRASTRUCTURE ONLY.

edt_prob and star_dist restate the *training-target* generators of the reference
(stardist/utils.py:edt_prob, stardist/geometry/geom2d.py:15-85 star_dist) closely enough to produce
realistic prob/dist maps the way tests/test_nms2D.py:70-73 does; they are not on the product path
and no parity claim is attached to them.
"""
import numpy as np
from scipy import ndimage as ndi


def ellipse_labels(shape, seed=0, fill=0.35, rmin=8, rmax=14):
    rng = np.random.default_rng(seed)
    H, W = shape
    lbl = np.zeros(shape, np.int32)
    target = fill * H * W
    filled, tries, k = 0, 0, 0
    m = int(np.ceil(rmax)) + 2
    yy, xx = np.mgrid[-m:m + 1, -m:m + 1]
    while filled < target and tries < 400000:
        tries += 1
        ry, rx = rng.uniform(rmin, rmax, 2)
        cy, cx = rng.integers(m, H - m), rng.integers(m, W - m)
        el = (yy / ry) ** 2 + (xx / rx) ** 2 <= 1
        sl = (slice(cy - m, cy + m + 1), slice(cx - m, cx + m + 1))
        if (lbl[sl][el] > 0).any():
            continue
        k += 1
        lbl[sl][el] = k
        filled += el.sum()
    return lbl


def image_from_labels(lbl, seed=0):
    rng = np.random.default_rng(seed + 1000)
    img = ndi.gaussian_filter((lbl > 0).astype(np.float32), 2) + rng.normal(0, 0.05, lbl.shape).astype(np.float32)
    lo, hi = np.percentile(img, 1), np.percentile(img, 99.8)
    return ((img - lo) / (hi - lo + 1e-20)).astype(np.float32)


def edt_prob(lbl):
    prob = np.zeros(lbl.shape, np.float32)
    objs = ndi.find_objects(lbl)
    for i, sl in enumerate(objs, 1):
        if sl is None: continue
        sl = tuple(slice(max(0, s.start - 1), s.stop + 1) for s in sl)
        m = lbl[sl] == i
        e = ndi.distance_transform_edt(m)
        prob[sl][m] = (e / (e.max() + 1e-10))[m]
    return prob


def star_dist(lbl, n_rays=32):
    """ray marching in unit steps until the label changes (geom2d.py:15-85 semantics incl. the
    half-step overshoot correction), vectorised over all foreground pixels"""
    H, W = lbl.shape
    dist = np.zeros((H, W, n_rays), np.float32)
    ys, xs = np.nonzero(lbl)
    val = lbl[ys, xs]
    for k in range(n_rays):
        phi = np.float32(2 * np.pi * k / n_rays)
        dy, dx = np.float32(np.sin(phi)), np.float32(np.cos(phi))   # row offset = d*sin(phi), col offset = d*cos(phi) (stardist2d.cpp:84-101, geom2d.py:141)
        y = np.zeros(len(ys), np.float32); x = np.zeros(len(ys), np.float32)
        alive = np.ones(len(ys), bool)
        res = np.zeros(len(ys), np.float32)
        for _ in range(4 * max(H, W)):
            if not alive.any(): break
            y[alive] += dy; x[alive] += dx
            ii = np.rint(ys + y).astype(int); jj = np.rint(xs + x).astype(int)
            out = alive & ((ii < 0) | (ii >= H) | (jj < 0) | (jj >= W))
            inb = alive & ~out
            diff = np.zeros_like(alive)
            diff[inb] = lbl[ii[inb], jj[inb]] != val[inb]
            stop = out | diff
            if stop.any():
                t_corr = np.float32(.5) / max(abs(dy), abs(dx))
                yy = y[stop] + (t_corr - 1) * dy; xx = x[stop] + (t_corr - 1) * dx
                res[stop] = np.sqrt(yy * yy + xx * xx)
                alive &= ~stop
        dist[ys, xs, k] = res
    return dist


def ellipsoid_labels(shape, seed=0, fill=0.30, rmin=5, rmax=9, aniso=(0.6, 1, 1)):
    """non-overlapping random ellipsoids (semi-axes U[rmin,rmax] * aniso), kept a margin away from the faces"""
    rng = np.random.default_rng(seed)
    lbl = np.zeros(shape, np.int32)
    target = fill * np.prod(shape)
    filled, k = 0, 0
    for _ in range(400000):
        if filled >= target: break
        r = rng.uniform(rmin, rmax, 3) * np.array(aniso)
        m = np.ceil(r).astype(int) + 1
        c = np.array([rng.integers(m[i], shape[i] - m[i]) for i in range(3)])
        sl = tuple(slice(c[i] - m[i], c[i] + m[i] + 1) for i in range(3))
        zz, yy, xx = np.mgrid[-m[0]:m[0] + 1, -m[1]:m[1] + 1, -m[2]:m[2] + 1]
        el = (zz / r[0]) ** 2 + (yy / r[1]) ** 2 + (xx / r[2]) ** 2 <= 1
        if (lbl[sl][el] > 0).any(): continue
        k += 1
        lbl[sl][el] = k
        filled += el.sum()
    return lbl


def star_dist3d(lbl, rays_vertices):
    """3-D star distances by unit-step ray marching along the (un-normalised) ray vertices until the label changes
    (stardist/lib/stardist3d_impl.cpp _COMMON_star_dist3D semantics incl. the half-step correction), vectorised over
    the foreground voxels; distances are in units of |vertex| like the reference's"""
    D, H, W = lbl.shape
    R = len(rays_vertices)
    zs, ys, xs = np.nonzero(lbl)
    val = lbl[zs, ys, xs]
    dist = np.zeros(lbl.shape + (R,), np.float32)
    for k in range(R):
        dz, dy, dx = (np.float32(v) for v in rays_vertices[k])
        z = np.zeros(len(zs), np.float32); y = np.zeros(len(zs), np.float32); x = np.zeros(len(zs), np.float32)
        alive = np.ones(len(zs), bool)
        res = np.zeros(len(zs), np.float32)
        for _ in range(4 * max(D, H, W)):
            if not alive.any(): break
            z[alive] += dz; y[alive] += dy; x[alive] += dx
            ii = np.rint(zs + z).astype(int); jj = np.rint(ys + y).astype(int); kk = np.rint(xs + x).astype(int)
            out = alive & ((ii < 0) | (ii >= D) | (jj < 0) | (jj >= H) | (kk < 0) | (kk >= W))
            inb = alive & ~out
            diff = np.zeros_like(alive)
            diff[inb] = lbl[ii[inb], jj[inb], kk[inb]] != val[inb]
            stop = out | diff
            if stop.any():
                res[stop] = np.sqrt(x[stop] ** 2 + y[stop] ** 2 + z[stop] ** 2) / np.float32(np.sqrt(dz * dz + dy * dy + dx * dx)) - np.float32(0.5)
                alive &= ~stop
        dist[zs, ys, xs, k] = np.maximum(res, 0)
    return dist


def volume_from_labels(lbl, seed=0):
    return image_from_labels(lbl, seed)


def calibrated_weights(config, seed=0, calib_shape=None, ridge=1e-3):
    """Glorot-uniform U-Net body (seeded) + heads fitted by ridge regression so that the network
    output on synthetic cell images resembles a trained StarDist (prob ~ edt_prob, dist ~ star_dist).
    A random-init net gives noise-like dist (negative -> clamped to 1e-3 -> degenerate polygons, no
    NMS work); this gives the benchmark a realistic candidate / instance structure while every
    layer of the real architecture still runs.  prob head is fitted in logit space."""
    from stardist_b200.models.weights import glorot_uniform_weights
    from . import unet_torch
    w = glorot_uniform_weights(config, seed=seed)
    if config.n_dim == 2:
        lbl = ellipse_labels(calib_shape or (512, 512), seed=seed + 17)
    else:
        lbl = ellipsoid_labels(calib_shape or (48, 128, 128), seed=seed + 17)
    img = image_from_labels(lbl, seed + 17)
    F = unet_torch.forward(config, w, img[None, ..., None], return_features=True)[0]      # [H,W,128]
    P = np.clip(edt_prob(lbl), 0.02, 0.98)
    if config.n_dim == 2:
        D = star_dist(lbl, config.n_rays)
    else:
        from stardist_b200.rays3d import rays_from_json
        D = star_dist3d(lbl, rays_from_json(config.rays_json).vertices)
    X = F.reshape(-1, F.shape[-1]).astype(np.float64)
    X1 = np.concatenate([X, np.ones((len(X), 1))], 1)
    Y = np.concatenate([np.log(P / (1 - P)).reshape(-1, 1), D.reshape(-1, config.n_rays)], 1).astype(np.float64)
    A = X1.T @ X1 + ridge * len(X) * np.eye(X1.shape[1]) * np.mean(np.diag(X1.T @ X1)) / len(X)
    B = X1.T @ Y
    sol = np.linalg.solve(A, B)
    k = tuple(config.unet_kernel_size)
    one = (1,) * config.n_dim
    w['prob'] = (sol[:-1, :1].reshape(one + (X.shape[1], 1)).astype(np.float32), sol[-1, :1].astype(np.float32))
    w['dist'] = (sol[:-1, 1:].reshape(one + (X.shape[1], config.n_rays)).astype(np.float32), sol[-1, 1:].astype(np.float32))
    return w
