"""torch-CPU fp32 restatement of the reference network (stand-in for TF/Keras + csbdeep, which
cannot be installed here).  TEST INFRASTRUCTURE ONLY.  [parity: pinned at whole-pipeline level for 2D by the
reference's own test numbers on its shipped 2D_demo weights (tests/test_cpu_oracle.py); UNPINNED at tensor level]

Follows model2d.py:310-349 / model3d.py:360-399 and csbdeep's unet_block topology:
conv(padding='same')+bias+ReLU, MaxPooling, UpSampling (nearest), Concatenate([up, skip]),
'features' conv, heads 1x1: prob sigmoid / dist linear.  Layer list shared with the product via
stardist_b200.models.weights.unet_layers (pure host metadata).
"""
import numpy as np
import torch
import torch.nn.functional as F


def forward(config, weights, x, dtype=torch.float32, return_features=False):
    """x: numpy [N, *spatial, Cin] channels-last -> (prob [N,*sp/g], dist [N,*sp/g,R]) numpy"""
    from stardist_b200.models.weights import unet_layers
    nd = config.n_dim
    conv = F.conv2d if nd == 2 else F.conv3d
    maxpool = F.max_pool2d if nd == 2 else F.max_pool3d
    t = torch.from_numpy(np.ascontiguousarray(x)).to(dtype)
    t = t.permute(0, nd + 1, *range(1, nd + 1)).contiguous()     # N C spatial

    def W(name):
        k, b = weights[name]
        kt = torch.from_numpy(k).to(dtype)
        kt = kt.permute(nd + 1, nd, *range(nd)).contiguous()     # (k.., Cin, Cout) -> (Cout, Cin, k..)
        return kt, torch.from_numpy(b).to(dtype)

    skips = {}
    for l in unet_layers(config):
        if l['kind'] == 'conv':
            kt, b = W(l['name'])
            t = conv(t, kt, b, padding='same')
            if l['act'] == 'relu': t = F.relu(t)
        elif l['kind'] == 'pool':
            if 'save_skip' in l: skips[l['save_skip']] = t
            t = maxpool(t, tuple(l['pool']))
        elif l['kind'] == 'up':
            t = F.interpolate(t, scale_factor=tuple(float(p) for p in l['pool']), mode='nearest')
            t = torch.cat([t, skips.pop(l['skip'])], dim=1)
        elif l['kind'] == 'head':
            break
    if return_features:
        return t.permute(0, *range(2, nd + 2), 1).contiguous().numpy()
    kp, bp = W('prob'); kd, bd = W('dist')
    prob = torch.sigmoid(conv(t, kp, bp))
    dist = conv(t, kd, bd)
    prob = prob[:, 0].numpy()
    dist = dist.permute(0, *range(2, nd + 2), 1).contiguous().numpy()
    return prob.astype(np.float32), dist.astype(np.float32)
