"""Device-side input preparation: percentile normalisation, the `scale=` zoom and the reflect padding.

Mirrors csbdeep.utils.normalize / csbdeep.data.PercentileNormalizer (call sites: stardist/scripts/predict2d.py:77,
stardist/models/base.py:398-404) and `ndi.zoom(img, scale, order=1)` + StarDistPadAndCropResizer.before
(stardist/models/base.py:725-735, :1162-1211) on tensors that are already in HBM (csrc/prep.cu).

Percentiles are numpy-exact: the device returns the two neighbouring ORDER STATISTICS of each percentile (radix select,
exact), and the interpolation between them is carried out by numpy's own routines on those two scalars with the dtypes
numpy.percentile would use for the original array -- so `mi` / `ma` carry the bits of np.percentile(x, p).
"""
import ctypes
import numpy as np
import torch
from . import _lib as L

# dtypes whose values survive the float32 upload unchanged (order statistics of the upload == those of the array)
EXACT_IN_F32 = (np.dtype(np.float32), np.dtype(np.float16), np.dtype(np.uint8), np.dtype(np.int8), np.dtype(np.uint16),
                np.dtype(np.int16), np.dtype(np.bool_))


def _virtual_index(n, p, dt):
    """numpy.percentile (method 'linear'): q = p / 100 in the array's float dtype, virtual index, neighbours, gamma"""
    from numpy.lib import _function_base_impl as F      # numpy >= 2
    q = np.asanyarray(np.true_divide(p, dt.type(100) if dt.kind == 'f' else 100))
    vi = np.asanyarray(F._QuantileMethods['linear']['get_virtual_index'](n, q))
    if np.issubdtype(vi.dtype, np.integer):
        return int(vi), int(vi), None
    prev, nxt = F._get_indexes(np.empty(n, np.bool_), vi, n)
    gamma = F._get_gamma(vi, prev, F._QuantileMethods['linear'])
    return int(prev) % n, int(nxt) % n, gamma          # numpy addresses the last element as -1 at the upper bound


def percentiles_device(x_dev, valid_shape, ps, src_dtype):
    """np.percentile(x, p) for every p in ps, x = the region [0, valid_shape) of the float32 device tensor x_dev (spatial
    shape, channel axis squeezed) that was uploaded from an array of dtype src_dtype.  Returns numpy scalars with the dtype
    numpy returns (float32 for float32 input, float64 for integer input, ...)."""
    from numpy.lib import _function_base_impl as F
    lib = L.require_cuda()
    dt = np.dtype(src_dtype)
    n = int(np.prod(valid_shape))
    plan = [_virtual_index(n, float(p), dt) for p in ps]
    ranks = []
    for lo, hi, _ in plan:
        ranks += [lo, hi]
    out = (ctypes.c_float * len(ranks))()
    shape = [int(s) for s in x_dev.shape]
    L.check(lib.sdb_select_ranks(L.ptr(x_dev), len(shape), L.iarr(shape), L.iarr(valid_shape), (ctypes.c_longlong * len(ranks))(*ranks),
                                len(ranks), out, L.stream_ptr()))
    vals = np.array(list(out), np.float32)
    res = []
    for k, (lo, hi, gamma) in enumerate(plan):
        a = vals[2 * k].astype(dt) if dt != np.bool_ else vals[2 * k] != 0
        b = vals[2 * k + 1].astype(dt) if dt != np.bool_ else vals[2 * k + 1] != 0
        res.append(np.asarray(a)[()] if gamma is None else F._lerp(np.asarray(a), np.asarray(b), gamma)[()])
    return res


def normalize_device(x_dev, valid_shape, pmin, pmax, src_dtype, clip=False, eps=1e-20, dtype=np.float32):
    """csbdeep.utils.normalize(x, pmin, pmax, clip=clip, eps=eps, dtype=float32) in place on the device tensor"""
    if dtype is not np.float32 and np.dtype(dtype) != np.float32:
        raise ValueError("normalize_device: only dtype=float32")
    lib = L.require_cuda()
    mi, ma = percentiles_device(x_dev, valid_shape, (pmin, pmax), src_dtype)
    mi32, ma32, eps32 = np.float32(mi), np.float32(ma), np.float32(eps)          # normalize_mi_ma: dtype(mi), dtype(ma), dtype(eps)
    den = np.float32(np.float32(ma32 - mi32) + eps32)
    L.check(lib.sdb_normalize_mi_ma(L.ptr(x_dev), x_dev.numel(), float(mi32), float(den), 1 if clip else 0, L.stream_ptr()))
    return mi, ma


def zoom_device(x_dev, zoom, src_dtype=np.float32):
    """scipy.ndimage.zoom(x, zoom, order=1) for a float32 device tensor of 2 or 3 dimensions that was uploaded from an
    array of dtype src_dtype: scipy returns the INPUT's dtype, i.e. integer images are rounded (and clipped) again"""
    lib = L.require_cuda()
    dt = np.dtype(src_dtype)
    if dt.kind in 'iu':
        rnd, lo, hi = 1, float(np.iinfo(dt).min), float(np.iinfo(dt).max)
    elif dt.kind == 'f':
        rnd, lo, hi = 0, 0.0, 0.0
    else:
        raise ValueError("zoom_device: unsupported source dtype %s" % dt)
    in_shape = [int(s) for s in x_dev.shape]
    out_shape = [int(round(s * z)) for s, z in zip(in_shape, zoom)]
    out = torch.empty(out_shape, dtype=torch.float32, device=x_dev.device)
    L.check(lib.sdb_zoom_linear(L.ptr(x_dev), len(in_shape), L.iarr(in_shape), L.iarr(out_shape), L.ptr(out), rnd, lo, hi, L.stream_ptr()))
    return out


def pad_reflect_end_device(x_dev, out_spatial):
    """np.pad(x, [(0, p)...], mode='reflect') on a channels-last device tensor [*spatial, C]"""
    lib = L.require_cuda()
    sp = [int(s) for s in x_dev.shape[:-1]]
    if list(out_spatial) == sp:
        return x_dev
    for s, o in zip(sp, out_spatial):
        if o - s > s - 1 and o != s:
            raise ValueError("reflect padding of %d needs an axis longer than %d" % (o - s, s))
    out = torch.empty(tuple(out_spatial) + (int(x_dev.shape[-1]),), dtype=torch.float32, device=x_dev.device)
    L.check(lib.sdb_pad_reflect_end(L.ptr(x_dev), len(sp), L.iarr(sp), L.iarr(list(out_spatial)), int(x_dev.shape[-1]), L.ptr(out), L.stream_ptr()))
    return out
