"""Freeze the reference's shipped `3D_demo` model (ResNet backbone) and its test volume as a fixture (run in the dev
container, where /root/reference exists):  python tests/golden/make_demo3d.py
  models/examples/3D_demo/{config.json, thresholds.json, weights_best.h5};  stardist/data/images/{img3d.tif, mask3d.tif}
-> tests/golden/demo3d.npz.  The reference's own test pins them to
   matching(mask, labels, thresh=0.5) -> (fp, tp, fn) == (0, 30, 21)   (tests/test_model3D.py:85-96)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from stardist_b200.io import h5lite
from PIL import Image

REF = "/root/reference"


def read_stack(path):
    im = Image.open(path)
    frames = []
    for i in range(getattr(im, "n_frames", 1)):
        im.seek(i); frames.append(np.array(im))
    return np.stack(frames)


w = h5lite.read_keras_weights(os.path.join(REF, "models/examples/3D_demo/weights_best.h5"))
out = {}
for name, (k, b) in w.items():
    out[name + "/kernel"] = k
    out[name + "/bias"] = b
out["config_json"] = np.frombuffer(open(os.path.join(REF, "models/examples/3D_demo/config.json"), "rb").read(), dtype=np.uint8)
out["thresholds_json"] = np.frombuffer(open(os.path.join(REF, "models/examples/3D_demo/thresholds.json"), "rb").read(), dtype=np.uint8)
out["img"] = read_stack(os.path.join(REF, "stardist/data/images/img3d.tif"))
out["mask"] = read_stack(os.path.join(REF, "stardist/data/images/mask3d.tif"))
path = os.path.join(ROOT, "tests", "golden", "demo3d.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes;", sum(v.size for k, v in out.items() if k.endswith("kernel") or k.endswith("bias")), "parameters")
