"""Compare the Qhull-free volume stages (geom3d.cuh / nms3d_pair.cuh, host build) with the reference's
qhull_overlap_kernel / qhull_overlap_convex_hulls (oracle/_ref/libsdref.so) on random polyhedron pairs."""
import ctypes, sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, "/root/reference"); sys.path.insert(0, ROOT)
ref = ctypes.CDLL(os.path.join(ROOT, "oracle/_ref/libsdref.so"))
hc = ctypes.CDLL(os.path.join(ROOT, "tests/hostcheck/_build/libhostcheck.so"))
for f in (ref.sdref_overlap_kernel, ref.sdref_overlap_convex, hc.hc_overlap_kernel, hc.hc_overlap_convex): f.restype = ctypes.c_float
P = ctypes.c_void_p

def rays(n):
    import importlib.util
    spec = importlib.util.spec_from_file_location("rays3d", "/root/reference/stardist/rays3d.py")
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    r = m.Rays_GoldenSpiral(n)
    return np.ascontiguousarray(r.vertices, np.float32), np.ascontiguousarray(r.faces, np.int32)

def run(n_pairs=2000, n_rays=96, radius=10, noise=.2, sep=12, seed=0):
    rng = np.random.default_rng(seed)
    verts, faces = rays(n_rays)
    out = []
    for _ in range(n_pairs):
        c1 = rng.uniform(20, 60, 3).astype(np.float32)
        c2 = (c1 + rng.uniform(-sep, sep, 3)).astype(np.float32)
        d1 = (radius * (1 + noise * rng.uniform(-1, 1, n_rays))).astype(np.float32)
        d2 = (radius * (1 + noise * rng.uniform(-1, 1, n_rays))).astype(np.float32)
        pv1 = (c1[None] + d1[:, None] * verts).astype(np.float32); pv2 = (c2[None] + d2[:, None] * verts).astype(np.float32)
        a = (P(pv1.ctypes.data), P(c1.ctypes.data), P(pv2.ctypes.data), P(c2.ctypes.data), P(faces.ctypes.data), n_rays, len(faces))
        rk = ref.sdref_overlap_kernel(*a); rc = ref.sdref_overlap_convex(*a)
        hk = hc.hc_overlap_kernel(*a); hcv = hc.hc_overlap_convex(*a[:4], n_rays)
        out.append((rk, hk, rc, hcv))
    return np.array(out, np.float64)

if __name__ == "__main__":
    args = [float(x) for x in sys.argv[1:]]
    o = run(*(int(args[0]), int(args[1])) if len(args) >= 2 else ())
    for name, a, b in (("kernel", o[:, 0], o[:, 1]), ("convex", o[:, 2], o[:, 3])):
        both = (a < 1e9) & (b < 1e9)
        rel = np.abs(a - b)[both] / np.maximum(1e-6, np.abs(a[both]))
        print(name, "n", len(a), "ref>0:", int((a > 0).sum()), "err-sentinel ref/ours:", int((a >= 1e9).sum()), int((b >= 1e9).sum()),
              "sentinel mismatch:", int(((a >= 1e9) != (b >= 1e9)).sum()), "zero mismatch:", int(((a == 0) != (b == 0)).sum()),
              "max rel diff:", rel.max() if len(rel) else None, "bit-equal floats:", int((a.astype(np.float32).view(np.int32) == b.astype(np.float32).view(np.int32)).sum()))
