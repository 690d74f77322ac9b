"""Loader of tests/golden/demo2d.npz (the reference's shipped 2D_demo model + test image, see make_demo2d.py)."""
import os, json
import numpy as np

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "demo2d.npz")
CONFIG_KEYS = ('n_rays', 'grid', 'n_channel_in', 'backbone', 'unet_n_depth', 'unet_kernel_size', 'unet_n_filter_base',
               'unet_n_conv_per_depth', 'unet_pool', 'unet_activation', 'unet_last_activation', 'unet_batch_norm',
               'net_conv_after_unet')
REFERENCE_TEST_STATS = (5, 114, 11)      # (fp, tp, fn), stardist tests/test_model2D.py:105


def load():
    """-> (config kwargs, weights dict, thresholds dict, img uint16 [512,512], mask uint16 [512,512])"""
    z = np.load(PATH)
    cfg = json.loads(bytes(z['config_json']).decode())
    thr = json.loads(bytes(z['thresholds_json']).decode())
    kwargs = {k: (tuple(v) if isinstance(v, list) else v) for k, v in cfg.items() if k in CONFIG_KEYS}
    names = sorted(k.rsplit('/', 1)[0] for k in z.files if k.endswith('/kernel'))
    weights = {n: (z[n + '/kernel'], z[n + '/bias']) for n in names}
    return kwargs, weights, thr, z['img'], z['mask']
