"""GPU tests of predict_instances_big (reference: tests/test_big.py:87-120 test_predict2D -- the result of
the block-wise prediction must equal the prediction on the whole image)."""
import os, sys
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _same_partition(a, b):
    """label maps equal up to a renaming of the ids"""
    if not np.array_equal(a > 0, b > 0): return False
    pairs = np.unique(np.stack([a[a > 0], b[a > 0]], 1), axis=0)
    return len(pairs) == len(np.unique(a[a > 0])) == len(np.unique(b[b > 0]))


def test_predict_instances_big_equals_whole_2d():
    import stardist_b200 as sd, bench_data
    cfg = sd.Config2D(n_rays=32)
    model = sd.StarDist2D(cfg, name=None, basedir=None, weights=bench_data.bench_weights_2d(cfg))
    img, _ = bench_data.synthetic_image((640, 704), seed=3)
    labels, polys = model.predict_instances(img)
    ctx = model._axes_tile_overlap('YX')
    assert all(0 < c < 128 for c in ctx)
    lb, pb = model.predict_instances_big(img, axes='YX', block_size=384, min_overlap=64, context=None, show_progress=False)
    assert lb.shape == labels.shape and lb.dtype == np.int32
    assert len(pb['prob']) == len(polys['prob']) > 50
    assert _same_partition(labels, lb)
    # same polygons (order differs: block by block) -- compare after lexsort, like the reference test
    def srt(p): 
        i = np.lexsort(tuple(p['points'].T)); return p['points'][i], p['prob'][i], p['coord'][i]
    for x, y in zip(srt(polys), srt(pb)):
        assert np.allclose(x, y, atol=1e-2)
