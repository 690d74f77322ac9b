"""Device executor of the U-Net forward pass: walks weights.unet_layers() and launches the CUDA
kernels of libstardist_b200.so on NHWC float32 tensors (torch is only the memory substrate).

Reference: `keras_model.predict(x[np.newaxis])` at stardist/models/base.py:408-410 for the graph
built in model2d.py:310-349.  The decoder's UpSampling+Concatenate is never materialised: the
conv kernel reads [upsample(x_lo), skip] through its loader address math.
"""
import numpy as np
import torch
from .. import _lib as L
from .weights import unet_layers


class UNetDevice2D:
    def __init__(self, config, weights, device=None):
        L.require_cuda()
        self.config = config
        self.device = torch.device("cuda") if device is None else torch.device(device)
        self.layers = unet_layers(config)
        self.w = {}
        for name, (k, b) in weights.items():
            self.w[name] = (torch.from_numpy(np.ascontiguousarray(k, dtype=np.float32)).to(self.device),
                            torch.from_numpy(np.ascontiguousarray(b, dtype=np.float32)).to(self.device))
        if config.unet_batch_norm:
            raise NotImplementedError("unet_batch_norm=True is not supported on this path")
        if tuple(config.unet_kernel_size) != (3, 3) or tuple(config.unet_pool) != (2, 2):
            raise NotImplementedError("only 3x3 kernels and 2x2 pooling are supported")
        if config.n_classes is not None:
            raise NotImplementedError("multi-class head is not supported yet")

    def _conv(self, x, x_lo, name, relu):
        lib = L.load()
        k, b = self.w[name]
        n, h, w, c_skip = x.shape
        c_lo = 0 if x_lo is None else x_lo.shape[-1]
        cout = k.shape[-1]
        assert k.shape[2] == c_skip + c_lo, (name, k.shape, c_skip, c_lo)
        out = torch.empty((n, h, w, cout), dtype=torch.float32, device=x.device)
        L.check(lib.sdb_conv3x3_2d(L.ptr(x), L.ptr(x_lo), n, h, w, c_skip, c_lo, L.ptr(k), L.ptr(b), cout,
                                  1 if relu else 0, L.ptr(out), L.stream_ptr()))
        return out

    def _pool(self, x):
        lib = L.load()
        n, h, w, c = x.shape
        out = torch.empty((n, h // 2, w // 2, c), dtype=torch.float32, device=x.device)
        L.check(lib.sdb_maxpool2x2_2d(L.ptr(x), n, h, w, c, L.ptr(out), L.stream_ptr()))
        return out

    def forward(self, x):
        """x: float32 [N,H,W,Cin] on the device -> (prob [N,H/g,W/g], dist [N,H/g,W/g,R])"""
        lib = L.load()
        assert x.dtype == torch.float32 and x.is_cuda and x.is_contiguous()
        skips = {}
        lo = None
        for l in self.layers:
            kind = l['kind']
            if kind == 'conv':
                act = l['act']
                if act not in ('relu', 'linear'):
                    raise NotImplementedError("activation %s" % act)
                if lo is not None:
                    x = self._conv(x, lo, l['name'], act == 'relu'); lo = None
                else:
                    x = self._conv(x, None, l['name'], act == 'relu')
            elif kind == 'pool':
                if tuple(l['pool']) != (2, 2):
                    raise NotImplementedError("anisotropic pooling")
                if 'save_skip' in l:
                    skips[l['save_skip']] = x
                x = self._pool(x)
            elif kind == 'up':
                lo = x                      # consumed by the next conv together with the skip
                x = skips.pop(l['skip'])
            elif kind == 'head':
                break
        feat = x
        n, h, w, cf = feat.shape
        R = self.config.n_rays
        (wp, bp), (wd, bd) = self.w['prob'], self.w['dist']
        prob = torch.empty((n, h, w), dtype=torch.float32, device=x.device)
        dist = torch.empty((n, h, w, R), dtype=torch.float32, device=x.device)
        L.check(lib.sdb_heads_2d(L.ptr(feat), n * h * w, cf, L.ptr(wp), L.ptr(bp), L.ptr(wd), L.ptr(bd), R,
                                L.ptr(prob), L.ptr(dist), L.stream_ptr()))
        return prob, dist
