"""tests/golden/big_cover.npz: the block covers of the reference's stardist/big.py (Block.cover / BlockND.cover, :171-280, :427-450)
for the parameter sets of the reference's own tests (tests/test_big.py:50-83) and a few more.  big.py is imported from
/root/reference with its third-party imports (skimage, csbdeep, the package-relative geometry import) replaced by stubs: the
cover arithmetic does not touch them.  Run in the build container: python tests/golden/make_big_cover.py"""
import importlib.util, json, os, sys, types
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from stardist_b200 import utils as U     # only axes_check_and_normalize / axes_dict / _raise for the stubbed csbdeep.utils

CASES_1D = ([(size, 4096, 128, 128, 16) for size in (7800, 7850, 7900, 7999, 8192, 4097, 20000)] +
            [(1024, 288, 32, 96, 8), (1000, 40, 11, 0, 1), (1000, 55, 13, 3, 3), (1000, 128, 20, 17, 6), (2000, 512, 50, 93, 6), (300, 300, 0, 0, 1), (301, 300, 10, 5, 4)])
CASES_ND = [((1024, 1024), 'YX', 288, 32, 96, 1), ((1040, 1392), 'YX', 256, 41, 80, 3), ((1040, 1392), 'YX', 128, 41, 17, 6),
            ((128, 512, 512), 'ZYX', (33, 71, 64), (9, 17, 17), 3, 1), ((128, 512, 512), 'ZYX', (62, 97, 93), (9, 17, 17), (0, 11, 9), 3),
            ((8192, 8192), 'YX', 2304, 128, 96, 8), ((512, 512, 512), 'ZYX', (160, 288, 288), (16, 32, 32), (16, 32, 32), (4, 4, 4))]


def load_reference_big():
    for name in ("skimage", "skimage.measure", "skimage.draw", "csbdeep", "csbdeep.utils", "stardist", "stardist.geometry"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["skimage.measure"].regionprops = None; sys.modules["skimage.draw"].polygon = None
    cu = sys.modules["csbdeep.utils"]; cu._raise = U._raise; cu.axes_check_and_normalize = U.axes_check_and_normalize; cu.axes_dict = U.axes_dict
    g = sys.modules["stardist.geometry"]; g.polygons_to_label_coord = None; g.polyhedron_to_label = None
    sys.modules["stardist"].__path__ = []
    spec = importlib.util.spec_from_file_location("stardist.big", "/root/reference/stardist/big.py")
    m = importlib.util.module_from_spec(spec); sys.modules["stardist.big"] = m; spec.loader.exec_module(m)
    return m


def describe_1d(blocks):
    return np.array([[b.start, b.end, b.slice_read.start, b.slice_read.stop, b.slice_crop_context.start, b.slice_crop_context.stop if b.slice_crop_context.stop is not None else 10 ** 9,
                      b.slice_write.start, b.slice_write.stop, int(b.at_begin), int(b.at_end)] for b in blocks], np.int64)


def describe_nd(blocks):
    rows = []
    for b in blocks:
        row = [b.id]
        for sl in (b.slice_read(), b.slice_crop_context(), b.slice_write()):
            for s in sl: row += [s.start if s.start is not None else 0, s.stop if s.stop is not None else 10 ** 9]
        rows.append(row)
    return np.array(rows, np.int64)


if __name__ == "__main__":
    big = load_reference_big()
    out = {"cases_1d": np.frombuffer(json.dumps(CASES_1D).encode(), np.uint8), "cases_nd": np.frombuffer(json.dumps(CASES_ND).encode(), np.uint8)}
    for i, (size, bs, mo, ctx, grid) in enumerate(CASES_1D):
        out["1d/%d" % i] = describe_1d(big.Block.cover(size, bs, mo, ctx, grid, verbose=False))
    # is_responsible (:89-122) on every interval (bmin, bmax) of a coarse lattice inside the block without context:
    # 1 = responsible, 0 = not, 2 = NotFullyVisible(False), 3 = NotFullyVisible(True)
    for i, (size, bs, mo, ctx, grid) in enumerate(CASES_1D):
        rows = []
        for k, b in enumerate(big.Block.cover(size, bs, mo, ctx, grid, verbose=False)):
            r_end = b.size - b.context_start - b.context_end
            pts = sorted(set(list(range(0, min(r_end, 40))) + list(range(max(0, r_end - 40), r_end + 1)) + list(range(0, r_end + 1, max(1, r_end // 23)))))
            for a in pts:
                for e in pts:
                    if not (0 <= a < e <= r_end): continue
                    try: v = int(bool(b.is_responsible((a, e))))
                    except big.NotFullyVisible as ex: v = 3 if ex.args[0] else 2
                    rows.append((k, a, e, v))
        out["resp/%d" % i] = np.array(rows, np.int32)
    for i, (shape, axes, bs, mo, ctx, grid) in enumerate(CASES_ND):
        out["nd/%d" % i] = describe_nd(big.BlockND.cover(shape, axes, bs, mo, ctx, grid))
    np.savez_compressed(os.path.join(HERE, "big_cover.npz"), **out)
    print({k: v.shape for k, v in out.items() if "/" in k})
