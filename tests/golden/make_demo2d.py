"""Freeze the reference's shipped `2D_demo` model and its test image as a fixture (run in the dev container, where
/root/reference exists):  python tests/golden/make_demo2d.py
  models/examples/2D_demo/{config.json, thresholds.json, weights_best.h5}  (read with stardist_b200.io.h5lite)
  stardist/data/images/{img2d.tif, mask2d.tif}                             (test_image_nuclei_2d, read with PIL)
-> tests/golden/demo2d.npz.  The reference's own test pins this model + image to
   matching(mask, labels, thresh=0.5) -> (fp, tp, fn) == (5, 114, 11)   (tests/test_model2D.py:92-106)."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from stardist_b200.io import h5lite
from PIL import Image

REF = "/root/reference"
w = h5lite.read_keras_weights(os.path.join(REF, "models/examples/2D_demo/weights_best.h5"))
out = {}
for name, (k, b) in w.items():
    out[name + "/kernel"] = k
    out[name + "/bias"] = b
out["config_json"] = np.frombuffer(open(os.path.join(REF, "models/examples/2D_demo/config.json"), "rb").read(), dtype=np.uint8)
out["thresholds_json"] = np.frombuffer(open(os.path.join(REF, "models/examples/2D_demo/thresholds.json"), "rb").read(), dtype=np.uint8)
out["img"] = np.array(Image.open(os.path.join(REF, "stardist/data/images/img2d.tif")))
out["mask"] = np.array(Image.open(os.path.join(REF, "stardist/data/images/mask2d.tif")))
path = os.path.join(ROOT, "tests", "golden", "demo2d.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes;", sum(v.size for k, v in out.items() if k.endswith("kernel") or k.endswith("bias")), "parameters")
