"""The 2D pair pre-filter (csrc/polyfast.cuh) must never change a decision: verify mode runs the closed-form
bound AND the exact Clipper-equivalent sweep on every pair and counts disagreements."""
import os, sys
import numpy as np, pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    import stardist_b200
    from stardist_b200 import _lib
    _lib.require_cuda()
    yield _lib
    _lib.nms2d_set_filter(1)


def _random_case(seed, n, n_rays, radius, noise, extent):
    rng = np.random.default_rng(seed)
    pts = rng.integers(0, extent, (n, 2)).astype(np.float32)
    base = radius * rng.uniform(0.6, 1.5, (n, 1))
    dist = (base * (1 + noise * rng.uniform(-1, 1, (n, n_rays)))).astype(np.float32)
    prob = rng.uniform(0.1, 1, n).astype(np.float32)
    o = np.argsort(prob, kind="stable")[::-1]
    return np.ascontiguousarray(dist[o]), np.ascontiguousarray(pts[o]), prob[o]


@pytest.mark.parametrize("seed,n,n_rays,radius,noise,extent", [
    (0, 20000, 32, 8, 0.15, 256), (1, 20000, 32, 3, 0.5, 128), (2, 6000, 64, 20, 0.3, 300),
    (3, 30000, 32, 12, 0.05, 400), (4, 5000, 16, 40, 0.6, 500), (5, 8000, 96, 10, 0.2, 200)])
@pytest.mark.parametrize("thr", [0.0, 0.4, 0.8])
def test_filter_modes_agree(L, seed, n, n_rays, radius, noise, extent, thr):
    from stardist_b200.lib.stardist2d import c_non_max_suppression_inds
    dist, pts, _ = _random_case(seed, n, n_rays, radius, noise, extent)
    res = {}
    for mode in (0, 1, 2):
        L.nms2d_set_filter(mode)
        L.nms2d_filter_stats(reset=True)
        res[mode] = c_non_max_suppression_inds(dist, pts, 1, 1, 0, np.float32(thr))
        st = L.nms2d_filter_stats(reset=True)
        if mode == 2:
            assert st["mismatches"] == 0, st
            assert st["pairs"] > 0
    assert np.array_equal(res[0], res[1]) and np.array_equal(res[0], res[2])


@pytest.mark.parametrize("name", list(cases.NMS2D_CASES))
def test_filter_verify_on_golden(L, name):
    from stardist_b200.lib.stardist2d import c_non_max_suppression_inds
    dist, pts, prob, thr = cases.nms2d_inputs(name)
    L.nms2d_set_filter(2)
    L.nms2d_filter_stats(reset=True)
    keep2 = c_non_max_suppression_inds(dist, pts, 1, 1, 0, np.float32(thr))
    st = L.nms2d_filter_stats(reset=True)
    assert st["mismatches"] == 0, st
    L.nms2d_set_filter(0)
    keep0 = c_non_max_suppression_inds(dist, pts, 1, 1, 0, np.float32(thr))
    assert np.array_equal(keep0, keep2)
