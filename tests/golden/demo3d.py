"""Loader of tests/golden/demo3d.npz (the reference's shipped 3D_demo ResNet model + test volume, see make_demo3d.py)."""
import os, json
import numpy as np

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "demo3d.npz")
CONFIG_KEYS = ('grid', 'anisotropy', 'backbone', 'n_channel_in', 'resnet_n_blocks', 'resnet_kernel_size', 'resnet_n_filter_base',
               'resnet_n_conv_per_block', 'resnet_activation', 'resnet_batch_norm', 'net_conv_after_resnet')
REFERENCE_TEST_STATS = (0, 30, 21)      # (fp, tp, fn), stardist tests/test_model3D.py:95


def load():
    """-> (rays_json, config kwargs, weights dict, thresholds dict, img uint16 [31,61,57], mask uint16)"""
    z = np.load(PATH)
    cfg = json.loads(bytes(z['config_json']).decode())
    thr = json.loads(bytes(z['thresholds_json']).decode())
    kwargs = {k: (tuple(v) if isinstance(v, list) else v) for k, v in cfg.items() if k in CONFIG_KEYS}
    names = sorted(k.rsplit('/', 1)[0] for k in z.files if k.endswith('/kernel'))
    weights = {n: (z[n + '/kernel'], z[n + '/bias']) for n in names}
    return cfg['rays_json'], kwargs, weights, thr, z['img'], z['mask']
