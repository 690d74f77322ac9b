"""Synthetic inputs and weights of the benchmark (pure host-side data generation; no oracle import).

Image: non-overlapping random ellipses, radius U[8,14] px, ~35 % fill, gaussian blur
sigma 2 + N(0, 0.05) noise, percentile-normalized float32.
Weights: seeded Glorot-uniform U-Net body of the named architecture (Keras default init) + prob/dist
1x1 heads fitted offline by ridge regression (tests/golden/make_bench_heads.py ->
tests/golden/bench_heads_2d.npz) so that the maps look like a trained StarDist's (prob ~ edt_prob,
dist ~ star_dist): a pure random-init net emits noise-like negative dist -> 1e-3-clamped degenerate
polygons and no NMS work, which would not be the reference's workload.  Every layer of the real
architecture runs; both bench arms use byte-identical weights."""
import os
import numpy as np
from scipy import ndimage as ndi

ROOT = os.path.dirname(os.path.abspath(__file__))


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


def synthetic_image(shape, seed=0):
    lbl = ellipse_labels(shape, seed=seed)
    rng = np.random.default_rng(seed + 1000)
    img = ndi.gaussian_filter((lbl > 0).astype(np.float32), 2) + rng.normal(0, 0.05, shape).astype(np.float32)
    lo, hi = np.percentile(img, 1), np.percentile(img, 99.8)
    return ((img - lo) / (hi - lo + 1e-20)).astype(np.float32), lbl


def bench_weights_2d(config, seed=0):
    from stardist_b200.models.weights import glorot_uniform_weights
    w = glorot_uniform_weights(config, seed=seed)
    h = np.load(os.path.join(ROOT, "tests", "golden", "bench_heads_2d.npz"))
    assert h['dist_kernel'].shape[-1] == config.n_rays and h['prob_kernel'].shape[-2] == config.net_conv_after_unet
    w['prob'] = (h['prob_kernel'], h['prob_bias'])
    w['dist'] = (h['dist_kernel'], h['dist_bias'])
    return w


# ------------------------------------------------------------------ 3-D (configs[2]: 128x512x512, Rays_GoldenSpiral 96)
CELL_3D = (64, 256, 256)


def ellipsoid_labels(shape, seed=0, fill=0.30, rmin=5, rmax=9, aniso=(0.6, 1, 1)):
    """non-overlapping random ellipsoids (semi-axes U[rmin,rmax] * aniso) with a margin to the faces"""
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


def synthetic_volume(shape, seed=0, cell=CELL_3D):
    """the SURVEY 8d volume: ellipsoids generated for one `cell` and tiled over `shape` (objects keep a margin from
    the cell faces, so the tiling is seamless), blurred + noisy + percentile-normalised like the 2-D image"""
    cell = tuple(min(c, s) for c, s in zip(cell, shape))
    assert all(s % c == 0 for s, c in zip(shape, cell))
    lbl_c = ellipsoid_labels(cell, seed=seed)
    rng = np.random.default_rng(seed + 1000)
    img_c = ndi.gaussian_filter((lbl_c > 0).astype(np.float32), 2)
    reps = tuple(s // c for s, c in zip(shape, cell))
    img = np.tile(img_c, reps) + rng.standard_normal(shape, dtype=np.float32) * np.float32(0.05)
    sub = img[::2, ::4, ::4]
    lo, hi = np.percentile(sub, 1), np.percentile(sub, 99.8)
    img -= np.float32(lo); img *= np.float32(1.0 / (hi - lo + 1e-20))
    return img, int(lbl_c.max()) * int(np.prod(reps))


def bench_config_3d(n_rays=96, anisotropy=None):
    from stardist_b200 import Config3D, Rays_GoldenSpiral
    return Config3D(rays=Rays_GoldenSpiral(n_rays, anisotropy=anisotropy))


def bench_weights_3d(config, seed=0):
    from stardist_b200.models.weights import glorot_uniform_weights
    w = glorot_uniform_weights(config, seed=seed)
    h = np.load(os.path.join(ROOT, "tests", "golden", "bench_heads_3d.npz"))
    assert h['dist_kernel'].shape[-1] == config.n_rays and h['prob_kernel'].shape[-2] == config.net_conv_after_unet
    w['prob'] = (h['prob_kernel'], h['prob_bias'])
    w['dist'] = (h['dist_kernel'], h['dist_bias'])
    return w


class TiledImage:
    """lazy (ny*H, nx*W[, ...]) periodic repetition of a tile: predict_instances_big only ever slices the blocks a rank
    owns, so no rank materialises (or uploads) the whole 8192^2 / 512^3 input"""
    def __init__(self, tile, reps):
        self.tile, self.reps = tile, tuple(reps)
        self.shape = tuple(s * r for s, r in zip(tile.shape, reps))
        self.ndim, self.dtype = tile.ndim, tile.dtype

    def __getitem__(self, sl):
        sl = sl if isinstance(sl, tuple) else (sl,)
        idx = []
        for s, n, t in zip(sl, self.shape, self.tile.shape):
            a, b, _ = s.indices(n)
            idx.append(np.arange(a, b) % t)
        return self.tile[np.ix_(*idx)]
