"""Deterministic input generators shared by make_golden.py and the tests.
Follow the reference's own test generators (tests/test_nms2D.py:9-16, tests/test_nms3D.py:8-15)
with the legacy seeded RandomState (np.random.seed(42), as tests/test_nms2D.py:83 does)."""
import numpy as np

NMS2D_CASES = {
    # name: (shape, radius, noise, n_rays, grid, prob_thresh, nms_thresh, seed)
    "r32_356x299": ((356, 299), 10, .1, 32, (1, 1), 0.9, 0.3, 42),
    "r11_114x217": ((114, 217), 10, .1, 11, (1, 1), 0.9, 0.3, 42),
    "r32_grid16": ((356, 299), 10, .1, 32, (16, 16), 0.3, 0.3, 42),
    "r32_noise0_thr0": ((200, 207), 10, 0.0, 32, (1, 1), 0.9, 0.0, 7),
    "r64_small": ((120, 130), 6, .3, 64, (1, 1), 0.9, 0.4, 3),
    "r32_tiny_radius": ((150, 150), 2, .5, 32, (1, 1), 0.9, 0.4, 5),
}


def create_random_data_2d(shape, radius, noise, n_rays, seed):
    rs = np.random.RandomState(seed)
    dist = radius * np.ones(shape + (n_rays,))
    noise = np.clip(noise, 0, 1)
    if noise > 0:
        dist *= (1 + noise * rs.uniform(-1, 1, dist.shape))
    prob = rs.uniform(0, 1, shape)
    return prob.astype(np.float32), dist.astype(np.float32)


def nms2d_inputs(name):
    """-> dist f32[n,R], points f32[n,2], scores f32[n] sorted by descending score (stable), nms_thresh"""
    shape, radius, noise, n_rays, grid, pthr, nthr, seed = NMS2D_CASES[name]
    prob, dist = create_random_data_2d(shape, radius, noise, n_rays, seed)
    prob = prob[::grid[0], ::grid[1]]; dist = dist[::grid[0], ::grid[1]]
    mask = prob > pthr
    b = 2
    m2 = np.zeros_like(mask); m2[b:-b, b:-b] = True
    mask &= m2
    points = np.stack(np.where(mask), axis=1)
    d = dist[mask]; s = prob[mask]
    ind = np.argsort(s, kind='stable')[::-1]
    d, s, points = d[ind], s[ind], points[ind]
    points = points * np.array(grid).reshape(1, 2)
    return np.ascontiguousarray(d, np.float32), np.ascontiguousarray(points, np.float32), s, np.float32(nthr)


# ---------------------------------------------------------------------------------- 3D
NMS3D_CASES = {
    # name: (shape, noise, n_rays, prob_thresh, nms_thresh, seed, anisotropy)   (tests/test_nms3D.py:25-30,46,62)
    "r5_thr0": ((33, 44, 55), 0.0, 5, 0.9, 0.0, 42, None),
    "r14_thr02": ((43, 31, 34), 0.0, 14, 0.9, 0.2, 42, None),
    "r22_thr04": ((33, 44, 55), 0.0, 22, 0.9, 0.4, 42, None),
    "r32_thr06": ((33, 44, 55), 0.0, 32, 0.9, 0.6, 42, None),
    "r32_noise01_thr01": ((33, 44, 55), 0.1, 32, 0.9, 0.1, 42, None),
    "r96_noise02_thr03": ((33, 44, 55), 0.2, 96, 0.9, 0.3, 7, None),
    "r96_aniso_thr03": ((24, 48, 50), 0.3, 96, 0.93, 0.3, 11, (2, 1, 1)),
}


def rays_golden_spiral(n, anisotropy=None):
    from stardist_b200.rays3d import Rays_GoldenSpiral     # host-only metadata, verified == reference rays3d.py
    return Rays_GoldenSpiral(n, anisotropy=anisotropy)


def create_random_data_3d(shape, noise, n_rays, seed):
    rs = np.random.RandomState(seed)
    dist = 10 * np.ones(shape + (n_rays,))
    noise = np.clip(noise, 0, 1)
    dist *= (1 + noise * rs.uniform(-1, 1, dist.shape))
    prob = rs.uniform(0, 1, shape)
    return prob.astype(np.float32), dist.astype(np.float32)


def nms3d_inputs(name):
    """-> dist f32[n,R], points f32[n,3], scores f32[n] (sorted desc, stable), rays, nms_thresh, shape"""
    shape, noise, n_rays, pthr, nthr, seed, aniso = NMS3D_CASES[name]
    prob, dist = create_random_data_3d(shape, noise, n_rays, seed)
    mask = prob > pthr
    b = 2
    m2 = np.zeros_like(mask); m2[b:-b, b:-b, b:-b] = True
    mask &= m2
    points = np.stack(np.where(mask), axis=1)
    d = dist[mask]; s = prob[mask]
    ind = np.argsort(s, kind='stable')[::-1]
    d, s, points = d[ind], s[ind], points[ind]
    return (np.ascontiguousarray(d, np.float32), np.ascontiguousarray(points, np.float32), np.ascontiguousarray(s, np.float32),
            rays_golden_spiral(n_rays, aniso), np.float32(nthr), shape)


# ---------------------------------------------------------------------------------- reference tests/test_nms3D.py:60-83
NMS3D_ACCURACY_CASES = [(noise, n_rays) for noise in (0, .2, .6, .9) for n_rays in (32, 65, 100)]


def nms3d_accuracy_inputs(noise, n_rays):
    """two polyhedra 3 voxels apart -> dist f64[2,R], points int[2,3], prob, rays, shape (test_nms_accuracy)"""
    rays = rays_golden_spiral(n_rays)
    dist = 10 * (1 + noise * np.sin(2 * np.pi * rays.vertices[:, :2].T))
    points = np.array([(20, 20, 20), (20, 20, 20 + 3)])
    return dist, points, np.array([1, .5]), rays, (40, 55, 66)
