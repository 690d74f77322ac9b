"""Generate the golden vectors under tests/golden/ by running the reference's own compiled C++
(oracle/_ref, built from /root/reference by `make -C oracle ref`) on the deterministic inputs of
cases.py.  Run in the build container:  python tests/golden/make_golden.py
For the 3D cases the reference is run with OMP_NUM_THREADS=1 (racy `anisotropy +=`,
stardist3d_impl.cpp:995-1010)."""
import os, sys, hashlib
os.environ.setdefault("OMP_NUM_THREADS", "1")
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
from oracle import ref_ext
import cases


def sha(*arrs):
    h = hashlib.sha256()
    for a in arrs: h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    s2 = ref_ext.stardist2d()
    out = {}
    for name in cases.NMS2D_CASES:
        d, p, s, thr = cases.nms2d_inputs(name)
        for kd in (1, 0):
            keep = s2.c_non_max_suppression_inds(d, p, kd, 1, 0, thr)
            out["%s/keep_kd%d" % (name, kd)] = np.packbits(keep)
        out[name + "/n"] = np.int64(len(d))
        out[name + "/sha"] = np.frombuffer(bytes.fromhex(sha(d, p)), np.uint8)
        print(name, len(d), int(keep.sum()))
    np.savez_compressed(os.path.join(HERE, "nms2d.npz"), **out)

    # ---- 3D: NMS keep masks + label volumes of the survivors (reference C++, single thread)
    s3 = ref_ext.stardist3d()
    out = {}
    for name in cases.NMS3D_CASES:
        d, p, s, rays, thr, shape = cases.nms3d_inputs(name)
        v = np.ascontiguousarray(rays.vertices, np.float32); f = np.ascontiguousarray(rays.faces, np.int32)
        keep = s3.c_non_max_suppression_inds(d, p, v, f, s, 1, 1, 0, thr)
        out[name + "/keep"] = np.packbits(keep)
        out[name + "/n"] = np.int64(len(d))
        dk, pk = d[keep], p[keep]
        labels = np.arange(1, len(dk) + 1, dtype=np.int32)
        for mode, mname in ((0, "full"), (1, "kernel"), (3, "bbox")):
            lbl = s3.c_polyhedron_to_label(dk, pk, v, f, labels, np.int32(mode), np.int32(0), np.int32(0), np.int32(0), shape)
            out["%s/label_%s" % (name, mname)] = lbl.astype(np.int32)
        lbl = s3.c_polyhedron_to_label(dk, pk, v, f, labels, np.int32(0), np.int32(0), np.int32(1), np.int32(-1), shape)
        out[name + "/label_full_overlap"] = lbl.astype(np.int32)
        print(name, len(d), int(keep.sum()), int((lbl != 0).sum()))
    np.savez_compressed(os.path.join(HERE, "nms3d.npz"), **out)

if __name__ == "__main__":
    main()
