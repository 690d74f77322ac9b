"""Measure |area_Clipper - I| for the fast overlap integral (polyfast.cuh, host build) against the reference
Clipper (oracle/_ref/libsdref.so), relative to polyfast's bound.  TEST TOOL ONLY.
Usage: python tests/tools/polyfast_fuzz.py [n_pairs] [n_rays] [radius] [noise] [seed] [span]"""
import ctypes, sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))
import clip_fuzz as cf

def fast(a, b):
    n_pairs, n = a.shape[:2]
    a32 = np.ascontiguousarray(a, np.int32); b32 = np.ascontiguousarray(b, np.int32)
    I = np.zeros(n_pairs); bd = np.zeros(n_pairs); K = np.zeros(n_pairs, np.int32)
    P = ctypes.c_void_p
    cf.hc.hc_fast_batch(P(a32.ctypes.data), P(b32.ctypes.data), n_pairs, n, P(I.ctypes.data), P(bd.ctypes.data), P(K.ctypes.data))
    return I, bd, K

def ref_area(a, b):
    n_pairs, n = a.shape[:2]
    a64 = np.ascontiguousarray(a, np.int64); b64 = np.ascontiguousarray(b, np.int64)
    r = np.zeros(n_pairs, np.float32)
    P = ctypes.c_void_p
    cf.ref.sdref_clip_area_batch(P(a64.ctypes.data), P(b64.ctypes.data), n_pairs, n, P(r.ctypes.data))
    return r

def shoelace(a):
    x = a[..., 0].astype(np.float64); y = a[..., 1].astype(np.float64)
    return 0.5 * np.abs((x * np.roll(y, -1, 1) - y * np.roll(x, -1, 1)).sum(1))

def report(a, b, tag, thr=0.4):
    r = ref_area(a, b).astype(np.float64)
    I, bd, K = fast(a, b)
    D = np.abs(r - I)
    ratio = D / bd
    w = int(np.argmax(ratio))
    amin = np.minimum(shoelace(a), shoelace(b)) + 1e-10
    decided = (I - bd > thr * amin * (1 + 1e-6)) | (I + bd < thr * amin * (1 - 1e-6))
    wrong = decided & (((I - bd > thr * amin * (1 + 1e-6)) & ~(r / amin > thr)) | ((I + bd < thr * amin * (1 - 1e-6)) & (r / amin > thr)))
    print(f"{tag}: pairs={len(r)} max|D|={D.max():.3f} max D/bound={ratio.max():.4f} (K={K[w]}, D={D[w]:.3f}, bound={bd[w]:.2f}) "
          f"mean bound={bd.mean():.2f} decided={decided.mean()*100:.1f}% wrong={int(wrong.sum())} I<-0.01: {int((I < -0.01).sum())}")
    return ratio.max()

if __name__ == "__main__":
    n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    n_rays = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    radius = float(sys.argv[3]) if len(sys.argv) > 3 else 10
    noise = float(sys.argv[4]) if len(sys.argv) > 4 else 0.1
    seed = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    span = float(sys.argv[6]) if len(sys.argv) > 6 else None
    a, b = cf.make_pairs(n_pairs, n_rays, radius, noise, seed, span)
    report(a, b, f"r={radius} noise={noise} rays={n_rays}")
