"""Network topology and weights of the U-Net backbone (host side).

Topology: stardist/models/model2d.py:310-349 / model3d.py:360-399 + csbdeep `unet_block`
(layer names and channel widths as evidenced by the shipped weight files):
  [grid stem: (conv x n_conv_per_depth @ base, maxpool) per grid doubling]
  down_level_{n}_no_{i}: base*2^n ; max_{n}
  middle_{i} (i < n_conv-1): base*2^depth ; middle_{n_conv}: base*2^max(0,depth-1)
  up: concat([upsample(x), skip_n]) -> up_level_{n}_no_{i} (i < n_conv-1): base*2^n ;
      up_level_{n}_no_{n_conv}: base*2^max(0,n-1)
  features: net_conv_after_unet ; prob: 1 (sigmoid) ; dist: n_rays (linear)
Kernels are in Keras layout (k..., Cin, Cout); Glorot-uniform init, zero bias (Keras defaults).
"""
import numpy as np


def unet_layers(config):
    """ordered list of layer dicts: name, kind ('conv'|'pool'|'up'), cin, cout, ..."""
    nd = config.n_dim
    k = tuple(config.unet_kernel_size)
    base, depth, nconv = config.unet_n_filter_base, config.unet_n_depth, config.unet_n_conv_per_depth
    pool = tuple(config.unet_pool)
    layers = []
    c = config.n_channel_in
    # grid stem
    pooled = np.array([1] * nd)
    si = 0
    while tuple(pooled) != tuple(config.grid):
        p = 1 + (np.asarray(config.grid) > pooled)
        pooled = pooled * p
        for _ in range(nconv):
            si += 1
            layers.append(dict(name='conv%dd_%d' % (nd, si), kind='conv', cin=c, cout=base, k=k, act=config.unet_activation))
            c = base
        layers.append(dict(name='stem_pool_%d' % si, kind='pool', pool=tuple(int(v) for v in p)))
    skips = []
    for n in range(depth):
        for i in range(nconv):
            layers.append(dict(name='down_level_%d_no_%d' % (n, i), kind='conv', cin=c, cout=base * 2 ** n, k=k, act=config.unet_activation))
            c = base * 2 ** n
        layers.append(dict(name='max_%d' % n, kind='pool', pool=pool, save_skip=n))
        skips.append(c)
    for i in range(nconv - 1):
        layers.append(dict(name='middle_%d' % i, kind='conv', cin=c, cout=base * 2 ** depth, k=k, act=config.unet_activation))
        c = base * 2 ** depth
    layers.append(dict(name='middle_%d' % nconv, kind='conv', cin=c, cout=base * 2 ** max(0, depth - 1), k=k, act=config.unet_activation))
    c = base * 2 ** max(0, depth - 1)
    for n in reversed(range(depth)):
        layers.append(dict(name='up_sampling_%d' % n, kind='up', pool=pool, skip=n))
        c_in = c + skips[n]
        for i in range(nconv - 1):
            layers.append(dict(name='up_level_%d_no_%d' % (n, i), kind='conv', cin=c_in, cout=base * 2 ** n, k=k,
                               act=config.unet_activation, cin_lo=(c if i == 0 else 0)))
            c_in = base * 2 ** n
        last_act = config.unet_last_activation if n == 0 else config.unet_activation
        layers.append(dict(name='up_level_%d_no_%d' % (n, nconv), kind='conv', cin=c_in, cout=base * 2 ** max(0, n - 1), k=k,
                           act=last_act, cin_lo=(c if nconv == 1 else 0)))
        c = base * 2 ** max(0, n - 1)
    c_base = c
    if config.net_conv_after_unet > 0:
        layers.append(dict(name='features', kind='conv', cin=c, cout=config.net_conv_after_unet, k=k, act=config.unet_activation))
        c = config.net_conv_after_unet
    layers.append(dict(name='prob', kind='head', cin=c, cout=1, k=(1,) * nd, act='sigmoid'))
    layers.append(dict(name='dist', kind='head', cin=c, cout=config.n_rays, k=(1,) * nd, act='linear'))
    layers += _class_layers(config, c_base, config.net_conv_after_unet, k, config.unet_activation)
    return layers


def _class_layers(config, c_base, n_after, k, act):
    """extra classification branch of multi-class models (model2d.py:339-347, model3d.py:388-396 / :436-444):
    features_class = conv(k)(backbone output) [if net_conv_after_* > 0], prob_class = softmax(conv 1^d, n_classes + 1).
    Kinds 'conv_class' / 'head_class' so that the single-class executors skip them."""
    if getattr(config, 'n_classes', None) is None:
        return []
    nd = config.n_dim
    out, c = [], c_base
    if n_after > 0:
        out.append(dict(name='features_class', kind='conv_class', cin=c_base, cout=int(n_after), k=tuple(k), act=act))
        c = int(n_after)
    out.append(dict(name='prob_class', kind='head_class', cin=c, cout=int(config.n_classes) + 1, k=(1,) * nd, act='softmax'))
    return out


def resnet_layers(config):
    """Layer list of the 3-D ResNet backbone (stardist/models/model3d.py:402-447 + csbdeep.internals.blocks.resnet_block):
      conv 7^3 (linear), conv 3^3 (linear), resnet_n_blocks x block(n_filter, pool), features conv 3^3 (+act), heads.
      block: x = conv(k, strides=pool) + act; (n_conv_per_block - 2) x [conv(k) + act]; conv(k);
             shortcut = conv(1^3, strides=pool)(inp) if pool > 1 or the channel count changes else inp;  out = act(shortcut + x)
    Keras auto-names the convolutions conv3d_1, conv3d_2, ... in creation order (the projection after the main path).
    entries: kind 'conv' (name, cin, cout, k, stride, act, src: 'cur' | 'block_in', dst: 'cur' | 'shortcut'),
             'block_begin', 'block_end' (act), 'head'."""
    nd = config.n_dim
    nd == 3 or _raise_value("the ResNet backbone exists for 3-D models only")
    if config.resnet_batch_norm:
        raise NotImplementedError("resnet_batch_norm=True is not supported on this path")
    k = tuple(int(v) for v in config.resnet_kernel_size)
    act = config.resnet_activation
    n_conv = int(config.resnet_n_conv_per_block)
    n_conv >= 2 or _raise_value("required: resnet_n_conv_per_block >= 2")
    layers, c, idx = [], config.n_channel_in, 0
    nf = int(config.resnet_n_filter_base)

    def conv(cin, cout, kk, stride=(1, 1, 1), a='linear', src='cur', dst='cur', name=None):
        nonlocal idx
        if name is None:
            idx += 1
            name = 'conv3d_%d' % idx
        layers.append(dict(name=name, kind='conv', cin=cin, cout=cout, k=tuple(kk), stride=tuple(int(v) for v in stride), act=a, src=src, dst=dst))

    conv(c, nf, (7, 7, 7)); conv(nf, nf, (3, 3, 3)); c = nf
    pooled = np.array([1, 1, 1])
    for _ in range(int(config.resnet_n_blocks)):
        pool = 1 + (np.asarray(config.grid) > pooled)
        pooled = pooled * pool
        if any(p > 1 for p in pool):
            nf *= 2
        layers.append(dict(name='block_begin', kind='block_begin'))
        conv(c, nf, k, stride=pool, a=act)
        for _ in range(n_conv - 2):
            conv(nf, nf, k, a=act)
        conv(nf, nf, k)
        if any(p != 1 for p in pool) or nf != c:
            conv(c, nf, (1, 1, 1), stride=pool, src='block_in', dst='shortcut')
        layers.append(dict(name='block_end', kind='block_end', act=act))
        c = nf
    c_base = c
    if config.net_conv_after_resnet > 0:
        conv(c, int(config.net_conv_after_resnet), k, a=act, name='features')
        c = int(config.net_conv_after_resnet)
    layers.append(dict(name='prob', kind='head', cin=c, cout=1, k=(1,) * nd, act='sigmoid'))
    layers.append(dict(name='dist', kind='head', cin=c, cout=config.n_rays, k=(1,) * nd, act='linear'))
    layers += _class_layers(config, c_base, int(config.net_conv_after_resnet), k, act)
    return layers


def _raise_value(msg):
    raise ValueError(msg)


def net_layers(config):
    """layer list of the configured backbone"""
    return resnet_layers(config) if getattr(config, 'backbone', 'unet') == 'resnet' else unet_layers(config)


def glorot_uniform_weights(config, seed=0):
    """dict name -> (kernel float32 (k..., Cin, Cout), bias float32 (Cout,)), seeded"""
    rng = np.random.default_rng(seed)
    w = {}
    for l in net_layers(config):
        if l['kind'] not in ('conv', 'head', 'conv_class', 'head_class'):
            continue
        k, cin, cout = l['k'], l['cin'], l['cout']
        rf = int(np.prod(k))
        limit = np.sqrt(6.0 / (rf * cin + rf * cout))
        kern = rng.uniform(-limit, limit, size=tuple(k) + (cin, cout)).astype(np.float32)
        w[l['name']] = (kern, np.zeros(cout, np.float32))
    return w


def count_params(weights):
    return int(sum(k.size + b.size for k, b in weights.values()))



def canonicalize_auto_names(config, weights):
    """Keras names un-named layers `conv2d`, `conv2d_1`, ... with a per-session counter (tf.keras 2.x starts without a
    suffix and keeps counting across models), and `load_weights` binds them by topology order, not by name.  The grid stem
    (model2d.py:316-325) and every ResNet convolution (csbdeep resnet_block) are such layers.  If the checkpoint's
    auto-named convolution groups do not carry exactly the names this architecture expects, they are re-bound
    positionally: sorted by their numeric suffix (creation order) and matched against the expected auto-named layers in
    topology order, with a kernel-shape check."""
    import re
    pat = re.compile(r'^conv%dd(?:_(\d+))?$' % config.n_dim)
    expected = [l for l in net_layers(config) if l['kind'] in ('conv', 'conv_class') and pat.match(l['name'])]
    if not expected or all(l['name'] in weights for l in expected):
        return weights
    have = sorted((k for k in weights if pat.match(k)), key=lambda k: int(pat.match(k).group(1) or 0))
    if len(have) != len(expected):
        raise ValueError("checkpoint holds %d auto-named convolution layers %s, the architecture needs %d"
                         % (len(have), have[:4], len(expected)))
    out = {k: v for k, v in weights.items() if not pat.match(k)}
    for l, k in zip(expected, have):
        want = tuple(l['k']) + (l['cin'], l['cout'])
        if tuple(np.asarray(weights[k][0]).shape) != want:
            raise ValueError("auto-named layer '%s' (position of '%s'): kernel shape %s, expected %s"
                             % (k, l['name'], tuple(np.asarray(weights[k][0]).shape), want))
        out[l['name']] = weights[k]
    return out
