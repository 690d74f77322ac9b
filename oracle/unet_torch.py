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


def _conv_same_tf(t, kt, bt, stride):
    """TensorFlow padding='same' for arbitrary strides: out = ceil(in/s), pad_total = max((out-1)*s + k - in, 0),
    pad_before = pad_total // 2 (the remainder goes behind) -- asymmetric for even inputs with stride 2."""
    pads = []
    for dim in reversed(range(3)):            # F.pad wants the last dimension first
        n_in, k, s = t.shape[2 + dim], kt.shape[2 + dim], stride[dim]
        n_out = -(-n_in // s)
        total = max((n_out - 1) * s + k - n_in, 0)
        pads += [total // 2, total - total // 2]
    return F.conv3d(F.pad(t, pads), kt, bt, stride=tuple(stride))


def forward_resnet(config, weights, x, dtype=torch.float32):
    """ResNet backbone (model3d.py:402-447 + csbdeep resnet_block): x numpy [N,D,H,W,Cin] -> (prob, dist) numpy"""
    from stardist_b200.models.weights import resnet_layers
    t = torch.from_numpy(np.ascontiguousarray(x)).to(dtype).permute(0, 4, 1, 2, 3).contiguous()

    def W(name):
        k, b = weights[name]
        return torch.from_numpy(k).to(dtype).permute(4, 3, 0, 1, 2).contiguous(), torch.from_numpy(b).to(dtype)

    block_in = shortcut = base = None
    for l in resnet_layers(config):
        if l['kind'] == 'block_begin':
            block_in, shortcut = t, None
        elif l['kind'] == 'conv':
            kt, bt = W(l['name'])
            if l['name'] == 'features': base = t
            y = _conv_same_tf(block_in if l['src'] == 'block_in' else t, kt, bt, l['stride'])
            if l['act'] == 'relu': y = F.relu(y)
            if l['dst'] == 'shortcut': shortcut = y
            else: t = y
        elif l['kind'] == 'block_end':
            t = (block_in if shortcut is None else shortcut) + t
            if l['act'] == 'relu': t = F.relu(t)
        elif l['kind'] == 'head':
            break
    kp, bp = W('prob'); kd, bd = W('dist')
    prob = torch.sigmoid(F.conv3d(t, kp, bp))[:, 0]
    dist = F.conv3d(t, kd, bd).permute(0, 2, 3, 4, 1)
    if getattr(config, 'n_classes', None) is not None:
        return prob.float().numpy(), dist.float().numpy(), _class_branch(config, W, F.conv3d, base, 3)
    return prob.float().numpy(), dist.float().numpy()


def forward(config, weights, x, dtype=torch.float32, return_features=False):
    """x: numpy [N, *spatial, Cin] channels-last -> (prob [N,*sp/g], dist [N,*sp/g,R]) numpy"""
    if getattr(config, 'backbone', 'unet') == 'resnet':
        return forward_resnet(config, weights, x, dtype=dtype)
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
    base = None
    for l in unet_layers(config):
        if l['kind'] == 'conv':
            kt, b = W(l['name'])
            if l['name'] == 'features': base = t
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
    odt = np.float64 if dtype == torch.float64 else np.float32       # float64 evaluations keep their precision
    if getattr(config, 'n_classes', None) is not None:
        return prob.astype(odt), dist.astype(odt), _class_branch(config, W, conv, t if base is None else base, nd)
    return prob.astype(odt), dist.astype(odt)


def _class_branch(config, W, conv, base, nd):
    """prob_class = softmax(conv1(features_class(backbone output)))  (model2d.py:339-347)"""
    t = base
    if 'features_class' in [l['name'] for l in __import__('stardist_b200.models.weights', fromlist=['net_layers']).net_layers(config)]:
        kt, b = W('features_class')
        t = F.relu(conv(t, kt, b, padding='same'))
    kc, bc = W('prob_class')
    pc = torch.softmax(conv(t, kc, bc), dim=1)
    return pc.permute(0, *range(2, nd + 2), 1).contiguous().numpy().astype(np.float32)
