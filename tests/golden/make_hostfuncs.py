"""tests/golden/hostfuncs.npz: outputs of the reference's OWN Python host functions on the path -- stardist/nms.py
(non_maximum_suppression[_3d][_sparse], :135-384), geometry/geom2d.py dist_to_coord (:130-146), geometry/geom3d.py
polyhedron_to_label (:100-198), matching.py matching / relabel_sequential (:109-232, :319-406) -- run from /root/reference
through tests/golden/_refpkg.py (third-party imports stubbed, compiled extensions = oracle/_ref) on the deterministic inputs
of cases.py.  The CPU tests hold the oracle restatements and the product's host mirrors against these.
Run in the build container: OMP_NUM_THREADS=1 python tests/golden/make_hostfuncs.py"""
import os, sys
os.environ.setdefault("OMP_NUM_THREADS", "1")
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refpkg, cases

NMS2D = ["r32_356x299", "r11_114x217", "r32_grid16", "r64_small"]
NMS3D = ["r14_thr02", "r32_noise01_thr01", "r96_aniso_thr03"]


def label_pair(seed, shape=(96, 120), n=40):
    rng = np.random.default_rng(seed)
    def one():
        lab = np.zeros(shape, np.int32)
        for i in range(1, n + 1):
            c = rng.integers(8, np.array(shape) - 8); r = rng.integers(3, 9)
            yy, xx = np.ogrid[:shape[0], :shape[1]]
            lab[(yy - c[0]) ** 2 + (xx - c[1]) ** 2 <= r * r] = i
        return lab
    a = one(); b = a.copy()
    b = np.roll(b, (rng.integers(-3, 4), rng.integers(-3, 4)), (0, 1)); b[b == 7] = 0; b[b == 11] = 55
    return a, b


if __name__ == "__main__":
    rnms = _refpkg.load("stardist.nms"); g2 = _refpkg.load("stardist.geometry.geom2d"); g3 = _refpkg.load("stardist.geometry.geom3d")
    rm = _refpkg.load("stardist.matching"); ru = _refpkg.load("stardist.utils")
    out = {}
    for name in NMS2D:
        shape, radius, noise, n_rays, grid, pthr, nthr, seed = cases.NMS2D_CASES[name]
        prob, dist = cases.create_random_data_2d(shape, radius, noise, n_rays, seed)
        prob = prob[::grid[0], ::grid[1]]; dist = dist[::grid[0], ::grid[1]]
        p, pr, d = rnms.non_maximum_suppression(dist, prob, grid=grid, b=2, nms_thresh=nthr, prob_thresh=pthr, verbose=False)
        out["nms2d/%s/points" % name] = p; out["nms2d/%s/prob" % name] = pr; out["nms2d/%s/dist" % name] = d
        mask = rnms._ind_prob_thresh(prob, pthr, b=2)
        pts = np.stack(np.where(mask), 1) * np.array(grid).reshape(1, 2)
        ps, prs, ds, inds = rnms.non_maximum_suppression_sparse(dist[mask], prob[mask], pts, nms_thresh=nthr, verbose=False)
        out["nms2d/%s/sparse_points" % name] = ps; out["nms2d/%s/sparse_inds" % name] = inds
        out["nms2d/%s/coord" % name] = g2.dist_to_coord(d, p, scale_dist=(1, 1)); out["nms2d/%s/coord_scaled" % name] = g2.dist_to_coord(d, p, scale_dist=(2, .5))
    for name in NMS3D:
        shape, noise, n_rays, pthr, nthr, seed, aniso = cases.NMS3D_CASES[name]
        prob, dist = cases.create_random_data_3d(shape, noise, n_rays, seed)
        rays = _refpkg.load("stardist.rays3d").Rays_GoldenSpiral(n_rays, anisotropy=aniso)
        p, pr, d = rnms.non_maximum_suppression_3d(dist, prob, rays, grid=(1, 1, 1), b=2, nms_thresh=nthr, prob_thresh=pthr, verbose=False)
        out["nms3d/%s/points" % name] = p; out["nms3d/%s/prob" % name] = pr; out["nms3d/%s/dist" % name] = d
        mask = rnms._ind_prob_thresh(prob, pthr, b=2)
        pts = np.stack(np.where(mask), 1)
        ps, prs, ds, inds = rnms.non_maximum_suppression_3d_sparse(dist[mask], prob[mask], pts, rays, nms_thresh=nthr, verbose=False)
        out["nms3d/%s/sparse_points" % name] = ps; out["nms3d/%s/sparse_inds" % name] = inds
        out["nms3d/%s/labels" % name] = g3.polyhedron_to_label(d, p, rays, shape, prob=pr, verbose=False).astype(np.int32)
    for seed in (0, 1, 2):
        a, b = label_pair(seed)
        out["match/%d/a" % seed] = a; out["match/%d/b" % seed] = b
        for crit in ("iou", "iot", "iop"):
            for thr in (0.3, 0.5, 0.9):
                m = rm.matching(a, b, thresh=thr, criterion=crit)
                out["match/%d/%s/%.1f" % (seed, crit, thr)] = np.array([m.fp, m.tp, m.fn, m.precision, m.recall, m.accuracy, m.f1, m.n_true, m.n_pred,
                                                                        m.mean_true_score, m.mean_matched_score, m.panoptic_quality], np.float64)
        rl, fw, inv = rm.relabel_sequential(b, offset=3)
        out["relabel/%d/out" % seed] = rl; out["relabel/%d/fw" % seed] = fw; out["relabel/%d/inv" % seed] = inv
    out["normalize_grid"] = np.array([ru._normalize_grid((2, 2, 2), 3), ru._normalize_grid([1, 2, 4], 3)], np.int64)
    np.savez_compressed(os.path.join(HERE, "hostfuncs.npz"), **out)
    print(len(out), "arrays,", os.path.getsize(os.path.join(HERE, "hostfuncs.npz")), "bytes")
