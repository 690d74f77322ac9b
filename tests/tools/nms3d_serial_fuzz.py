"""Serial host restatement of the GPU 3-D NMS (tests/hostcheck hc_nms3d_serial: the arithmetic of geom3d.cuh / nms3d_pair.cuh
in the reference's greedy order) against the reference's c_non_max_suppression_inds (oracle/_ref) on random candidate clouds,
for ray counts, anisotropies and flags beyond the committed golden cases.  Run with OMP_NUM_THREADS=1 (the reference's
anisotropy sum is racy).  Prints one JSON line.  CPU only."""
import ctypes, json, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases
from oracle import ref_ext

subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "hostcheck")], check=True, stdout=subprocess.DEVNULL)
hc = ctypes.CDLL(os.path.join(ROOT, "tests", "hostcheck", "_build", "libhostcheck.so"))
P = ctypes.c_void_p
hc.hc_nms3d_serial.argtypes = [P, P, P, P, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int, P, P]

CASES = [  # shape, noise, n_rays, prob_thresh, nms_thresh, seed, anisotropy, use_bbox, use_kdtree
    ((20, 26, 30), 0.2, 65, 0.93, 0.3, 1, None, 1, 1), ((20, 26, 30), 0.4, 100, 0.94, 0.2, 2, None, 1, 1),
    ((18, 24, 26), 0.3, 187, 0.96, 0.4, 3, None, 1, 1), ((24, 30, 33), 0.1, 32, 0.9, 0.1, 4, (2, 1, 1), 1, 1),
    ((24, 30, 33), 0.6, 48, 0.92, 0.5, 5, (1, 1.5, 3), 1, 1), ((22, 28, 31), 0.2, 24, 0.9, 0.3, 6, None, 0, 1),
    ((22, 28, 31), 0.2, 24, 0.9, 0.3, 7, None, 1, 0), ((22, 28, 31), 0.9, 40, 0.93, 0.05, 8, None, 1, 1),
    ((30, 34, 36), 0.0, 16, 0.9, 0.25, 9, None, 1, 1), ((16, 40, 44), 0.5, 70, 0.93, 0.6, 10, (4, 1, 1), 1, 1),
]
MINI = [  # seconds, for the pytest run (tests/test_cpu_oracle.py)
    ((12, 16, 18), 0.3, 65, 0.93, 0.3, 1, None, 1, 1), ((12, 16, 18), 0.4, 100, 0.94, 0.2, 2, None, 1, 1),
    ((14, 18, 20), 0.2, 24, 0.92, 0.3, 6, (2, 1, 1), 1, 1), ((14, 18, 20), 0.5, 32, 0.92, 0.4, 7, None, 0, 0),
]
mode = sys.argv[1] if len(sys.argv) > 1 else "full"
out = dict(cases=0, candidates=0, mismatches=0, stage_counts=[0] * 5, per_case=[])
for shape, noise, n_rays, pthr, nthr, seed, aniso, use_bbox, use_kd in (MINI if mode == "mini" else CASES):
    prob, dist = cases.create_random_data_3d(shape, noise, n_rays, seed)
    mask = prob > pthr
    m2 = np.zeros_like(mask); m2[2:-2, 2:-2, 2:-2] = True
    mask &= m2
    points = np.stack(np.where(mask), axis=1)
    d = dist[mask]; s = prob[mask]
    ind = np.argsort(s, kind='stable')[::-1]
    d = np.ascontiguousarray(d[ind], np.float32); p = np.ascontiguousarray(points[ind], np.float32); s = np.ascontiguousarray(s[ind], np.float32)
    rays = cases.rays_golden_spiral(n_rays, aniso)
    v = np.ascontiguousarray(rays.vertices, np.float32); f = np.ascontiguousarray(rays.faces, np.int32)
    want = ref_ext.stardist3d().c_non_max_suppression_inds(d, p, v, f, s, int(use_bbox), int(use_kd), 0, np.float32(nthr))
    res = {}
    for variant in (0, 1):
        keep = np.zeros(len(d), np.uint8); sc = np.zeros(5, np.int32)
        hc.hc_nms3d_serial(d.ctypes.data, p.ctypes.data, v.ctypes.data, f.ctypes.data, len(d), n_rays, len(f), ctypes.c_float(nthr),
                           int(use_bbox), int(use_kd), variant, keep.ctypes.data, sc.ctypes.data)
        res[variant] = (keep.astype(bool), sc)
    mm = int(np.count_nonzero(res[0][0] != want)) + int(np.count_nonzero(res[1][0] != want))
    out['cases'] += 1; out['candidates'] += len(d); out['mismatches'] += mm
    out['stage_counts'] = [int(a + b) for a, b in zip(out['stage_counts'], res[0][1])]
    out['per_case'].append(dict(n_rays=n_rays, n=len(d), kept=int(want.sum()), mismatches=mm, stages=[int(x) for x in res[0][1]]))
print(json.dumps(out))
