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


def test_device_block_pipeline_equals_host_pipeline_2d(monkeypatch):
    """the device-resident block pipeline (sdb_label_bbox / remap / write) reproduces the host pipeline
    (BlockND.filter_objects + relabel_sequential + BlockND.write, pinned against the reference's big.py) bit for bit"""
    import stardist_b200 as sd, bench_data
    cfg = sd.Config2D(n_rays=32)
    model = sd.StarDist2D(cfg, name=None, basedir=None, weights=bench_data.bench_weights_2d(cfg))
    img, _ = bench_data.synthetic_image((1024, 896), seed=5)
    kw = dict(axes='YX', block_size=384, min_overlap=64, context=48, show_progress=False)
    monkeypatch.setenv("STARDIST_B200_BIG", "host")
    lh, ph = model.predict_instances_big(img, **kw)
    monkeypatch.setenv("STARDIST_B200_BIG", "device")
    ld, pd_ = model.predict_instances_big(img, **kw)
    assert ld.dtype == lh.dtype and np.array_equal(ld, lh)
    assert set(pd_) == set(ph)
    for k in ph:
        assert np.array_equal(np.asarray(pd_[k]), np.asarray(ph[k])), k
    # labels_out given by the caller
    out = np.zeros(img.shape, np.int32)
    l2, _ = model.predict_instances_big(img, labels_out=out, **kw)
    assert l2 is out and np.array_equal(out, lh)


def test_device_block_pipeline_equals_host_pipeline_3d(monkeypatch):
    import stardist_b200 as sd, bench_data
    cfg = bench_data.bench_config_3d(96)
    model = sd.StarDist3D(cfg, name=None, basedir=None, weights=bench_data.bench_weights_3d(cfg))
    vol, _ = bench_data.synthetic_volume((64, 128, 128), seed=2, cell=(64, 128, 128))
    kw = dict(axes='ZYX', block_size=(48, 96, 96), min_overlap=(16, 24, 24), context=(8, 16, 16), show_progress=False,
              prob_thresh=0.7, nms_thresh=0.3)
    monkeypatch.setenv("STARDIST_B200_BIG", "host")
    lh, ph = model.predict_instances_big(vol, **kw)
    monkeypatch.setenv("STARDIST_B200_BIG", "device")
    ld, pd_ = model.predict_instances_big(vol, **kw)
    assert len(ph['prob']) > 20
    assert np.array_equal(ld, lh)
    for k in ('prob', 'points', 'dist'):
        assert np.array_equal(pd_[k], ph[k]), k


def test_label_bbox_kernel_equals_find_objects():
    """sdb_label_bbox == scipy.ndimage.find_objects (== skimage regionprops bbox, big.py:373) incl. absent labels"""
    import torch
    from scipy import ndimage as ndi
    from stardist_b200 import _lib as L
    lib = L.require_cuda()
    rng = np.random.default_rng(0)
    for shape in ((257, 301), (33, 70, 45)):
        lab = np.zeros(shape, np.int32)
        nlab = 60
        for i in range(1, nlab + 1):
            if i % 7 == 0: continue                        # absent labels
            c = [rng.integers(0, s) for s in shape]
            r = [rng.integers(1, 12) for _ in shape]
            sl = tuple(slice(max(0, a - b), a + b) for a, b in zip(c, r))
            m = rng.random(lab[sl].shape) < 0.6
            lab[sl][m] = i
        t = torch.from_numpy(lab).cuda()
        bb = torch.empty((nlab + 1) * 6 + 1, dtype=torch.int32, device='cuda')
        L.check(lib.sdb_label_bbox(L.ptr(t), t.dim(), L.iarr(t.shape), nlab, L.ptr(bb), L.ptr(bb[(nlab + 1) * 6:]), L.stream_ptr()))
        bb = bb.cpu().numpy()
        assert bb[-1] == 0
        bb = bb[:-1].reshape(nlab + 1, 6)
        nd = len(shape)
        for i, sl in enumerate(ndi.find_objects(lab, max_label=nlab), 1):
            lo, hi = bb[i, 3 - nd:3], bb[i, 6 - nd:6]
            if sl is None:
                assert (hi < lo).all()
            else:
                assert tuple(lo) == tuple(s.start for s in sl) and tuple(hi + 1) == tuple(s.stop for s in sl), (i, lo, hi, sl)
