"""tests/golden/big_blocks.npz: the reference's BlockND pipeline (stardist/big.py: read :312, crop_context :316, filter_objects
:340-413, write :319-326, translate_coordinates :415-425) run block by block on synthetic label images, with per-object
polygon dictionaries, through tests/golden/_refpkg.py (skimage.measure.regionprops replaced by a scipy.ndimage.find_objects
stand-in that provides label / bbox / image).  Stored per case: the label image, for every block the filtered labels and
the surviving object ids / translated points, and the re-assembled image.
Run in the build container: python tests/golden/make_big_blocks.py"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refpkg

CASES = [  # shape, axes, block_size, min_overlap, context, grid, n objects, max radius, seed
    ((300, 420), 'YX', 96, 24, 8, 1, 120, 9, 0), ((257, 390), 'YX', (128, 100), (24, 26), (16, 4), (4, 2), 90, 10, 1),
    ((40, 96, 110), 'ZYX', (24, 48, 56), (10, 16, 16), (2, 8, 4), (1, 2, 2), 60, 6, 2),
]


def label_image(shape, n, rmax, seed):
    rng = np.random.default_rng(seed)
    lab = np.zeros(shape, np.int32)
    grids = np.ogrid[tuple(slice(0, s) for s in shape)]
    k = 0
    for _ in range(20 * n):
        if k >= n: break
        r = rng.integers(2, rmax + 1)
        c = [rng.integers(0, s) for s in shape]
        m = sum(((g - ci) / (r if i else max(1, r // (2 if len(shape) == 3 else 1)))) ** 2 for i, (g, ci) in enumerate(zip(grids, c))) <= 1
        if lab[m].any(): continue
        k += 1; lab[m] = k
    return lab


def block_inputs(block, lab, axes):
    """what a model would return for the block: labels of the read region renumbered 1..n, polys with points / prob / dist"""
    x = block.read(lab, axes=axes)
    ids = np.unique(x); ids = ids[ids > 0]
    local = np.zeros_like(x)
    for j, i in enumerate(ids, 1): local[x == i] = j
    pts = np.array([np.round(np.mean(np.argwhere(local == j), 0)) for j in range(1, len(ids) + 1)]).reshape(len(ids), x.ndim)
    polys = dict(points=pts, prob=(ids % 97 / 97.0).astype(np.float32), dist=np.tile(ids[:, None].astype(np.float32), (1, 5)), rays_faces=np.arange(6))
    return local, polys, ids


if __name__ == "__main__":
    big = _refpkg.load("stardist.big")
    out = {}
    for ci, (shape, axes, bs, mo, ctx, grid, n, rmax, seed) in enumerate(CASES):
        lab = label_image(shape, n, rmax, seed)
        out["%d/label" % ci] = lab
        blocks = big.BlockND.cover(shape, axes, bs, mo, ctx, grid)
        result = np.zeros_like(lab); offset = 1
        for b in blocks:
            local, polys, ids = block_inputs(b, lab, axes)
            cropped = b.crop_context(local, axes=axes)
            kept, polys_out = b.filter_objects(cropped, polys, axes=axes)
            out["%d/%d/kept" % (ci, b.id)] = kept
            out["%d/%d/points" % (ci, b.id)] = polys_out["points"]; out["%d/%d/prob" % (ci, b.id)] = polys_out["prob"]; out["%d/%d/dist" % (ci, b.id)] = polys_out["dist"]
            b.write(result, np.where(kept > 0, b.crop_context(b.read(lab, axes=axes), axes=axes), 0), axes=axes)      # global ids of the kept objects
        out["%d/reassembled" % ci] = result
        print(ci, shape, len(blocks), "blocks, objects", int(lab.max()), "reassembled == label:", bool(np.array_equal(result, lab)))
    np.savez_compressed(os.path.join(HERE, "big_blocks.npz"), **out)
    print(os.path.getsize(os.path.join(HERE, "big_blocks.npz")), "bytes")


# ------------------------------------------------------------------------------------------------------------------------
# predict_instances_big (stardist/models/base.py:838-983) on a stand-in model: `img` IS a label image, predict_instances returns
# the objects visible in the block (renumbered 1..n) with their polys -- the bookkeeping under test is the reference's / ours.
class FakeModel:
    def __init__(self, axes_out, grid, overlap):
        import types
        self._axes_out = axes_out + 'C'; self.config = types.SimpleNamespace(axes=axes_out + 'C')
        self._grid = dict(zip(axes_out, grid)); self._overlap = dict(zip(axes_out, overlap))

    def _axes_div_by(self, axes): return tuple(self._grid.get(a, 1) for a in axes)
    def _axes_tile_overlap(self, axes): return tuple(self._overlap.get(a, 0) for a in axes)

    def predict_instances(self, x, axes=None, **kwargs):
        ids = np.unique(x); ids = ids[ids > 0]
        local = np.zeros(x.shape, np.int32)
        for j, i in enumerate(ids, 1): local[x == i] = j
        pts = np.array([np.round(np.mean(np.argwhere(local == j), 0)) for j in range(1, len(ids) + 1)]).reshape(len(ids), x.ndim)
        polys = dict(points=pts, prob=(ids % 97 / 97.0).astype(np.float32), dist=np.tile(ids[:, None].astype(np.float32), (1, 5)), rays_faces=np.arange(6))
        return local, polys


BIG_CASES = [(0, dict(block_size=96, min_overlap=24, context=8), (1, 1), (6, 6)), (1, dict(block_size=(128, 100), min_overlap=(24, 26), context=None), (4, 2), (16, 4)),
             (2, dict(block_size=(24, 48, 56), min_overlap=(10, 16, 16), context=(2, 8, 4)), (1, 2, 2), (2, 8, 4))]


def run_big(method, label, axes, kw, grid, overlap):
    m = FakeModel(axes, grid, overlap)
    return method(m, label, axes=axes, show_progress=False, **{k: v for k, v in kw.items()})


if __name__ == "__main__":
    m2, _ = _refpkg.load_models()
    base = sys.modules["stardist.models.base"]
    g = dict(np.load(os.path.join(HERE, "big_blocks.npz")))
    for ci, kw, grid, overlap in BIG_CASES:
        shape, axes = CASES[ci][0], CASES[ci][1]
        labels_out, polys = run_big(base.StarDistBase.predict_instances_big, g["%d/label" % ci], axes, kw, grid, overlap)
        g["big/%d/labels" % ci] = labels_out
        for k in ("points", "prob", "dist", "rays_faces"): g["big/%d/%s" % (ci, k)] = polys[k]
        print("predict_instances_big", ci, labels_out.dtype, int(labels_out.max()), len(polys["prob"]))
    np.savez_compressed(os.path.join(HERE, "big_blocks.npz"), **g)
