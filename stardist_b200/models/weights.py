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
    if config.net_conv_after_unet > 0:
        layers.append(dict(name='features', kind='conv', cin=c, cout=config.net_conv_after_unet, k=k, act=config.unet_activation))
        c = config.net_conv_after_unet
    layers.append(dict(name='prob', kind='head', cin=c, cout=1, k=(1,) * nd, act='sigmoid'))
    layers.append(dict(name='dist', kind='head', cin=c, cout=config.n_rays, k=(1,) * nd, act='linear'))
    return layers


def glorot_uniform_weights(config, seed=0):
    """dict name -> (kernel float32 (k..., Cin, Cout), bias float32 (Cout,)), seeded"""
    rng = np.random.default_rng(seed)
    w = {}
    for l in unet_layers(config):
        if l['kind'] not in ('conv', 'head'):
            continue
        k, cin, cout = l['k'], l['cin'], l['cout']
        rf = int(np.prod(k))
        limit = np.sqrt(6.0 / (rf * cin + rf * cout))
        kern = rng.uniform(-limit, limit, size=tuple(k) + (cin, cout)).astype(np.float32)
        w[l['name']] = (kern, np.zeros(cout, np.float32))
    return w


def count_params(weights):
    return int(sum(k.size + b.size for k, b in weights.values()))
