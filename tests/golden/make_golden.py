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

if __name__ == "__main__":
    main()
