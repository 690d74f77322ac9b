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
