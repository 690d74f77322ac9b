"""Host-side helpers that the prediction path needs.

Reference: stardist/utils.py:54-77 (_is_power_of_2, _normalize_grid); csbdeep.utils
(axes_check_and_normalize, axes_dict, normalize -- csbdeep is an un-vendored dependency of the
reference, setup.py:140; behaviour restated from its documented semantics).
"""
import numpy as np
from collections.abc import Iterable


def _raise(e):
    raise e


def _is_power_of_2(i):
    assert i > 0
    e = np.log2(i)
    return e == int(e)


def _normalize_grid(grid, n):
    # stardist/utils.py:60-77
    try:
        grid = tuple(grid)
        (len(grid) == n and all(map(np.isscalar, grid)) and all(map(_is_power_of_2, grid))) or _raise(TypeError())
        return tuple(int(g) for g in grid)
    except (TypeError, AssertionError):
        raise ValueError("grid = {grid} must be a list/tuple of length {n} with values that are power of 2".format(grid=grid, n=n))


def axes_check_and_normalize(axes, length=None, disallowed=None, return_allowed=False):
    """S(ample), T(ime), C(hannel), Z, Y, X"""
    allowed = 'STCZYX'
    axes is not None or _raise(ValueError('axis cannot be None.'))
    axes = str(axes).upper()
    for a in axes:
        a in allowed or _raise(ValueError("invalid axis '%s', must be one of %s." % (a, list(allowed))))
        (disallowed is None or a not in disallowed) or _raise(ValueError("disallowed axis '%s'." % a))
        axes.count(a) == 1 or _raise(ValueError("axis '%s' occurs more than once." % a))
    (length is None or len(axes) == length) or _raise(ValueError('axes (%s) must be of length %d.' % (axes, length)))
    return (axes, allowed) if return_allowed else axes


def axes_dict(axes):
    axes, allowed = axes_check_and_normalize(axes, return_allowed=True)
    return {a: None if axes.find(a) == -1 else axes.find(a) for a in allowed}


def move_image_axes(x, fr, to, adjust_singletons=False):
    fr = axes_check_and_normalize(fr, length=x.ndim)
    to = axes_check_and_normalize(to)
    fr_initial, x_shape_initial = fr, x.shape
    adjust_singletons = bool(adjust_singletons)
    if adjust_singletons:
        # remove axes not present in 'to'
        slices = [slice(None) for _ in x.shape]
        for i, a in enumerate(fr):
            if (a not in to) and (x.shape[i] == 1):
                slices[i] = 0
                fr = fr.replace(a, '')
        x = x[tuple(slices)]
        # add dummy axes present in 'to'
        for i, a in enumerate(to):
            if a not in fr:
                x = np.expand_dims(x, -1)
                fr += a
    if set(fr) != set(to):
        _adjusted = '(adjusted to %s and %s) ' % (x.shape, fr) if adjust_singletons else ''
        raise ValueError('image with shape %s and axes %s %snot compatible with target axes %s.'
                         % (x_shape_initial, fr_initial, _adjusted, to))
    ax_from, ax_to = axes_dict(fr), axes_dict(to)
    if fr == to:
        return x
    return np.moveaxis(x, [ax_from[a] for a in fr], [ax_to[a] for a in fr])


def normalize(x, pmin=3, pmax=99.8, axis=None, clip=False, eps=1e-20, dtype=np.float32):
    """Percentile-based image normalization (csbdeep.utils.normalize)."""
    mi = np.percentile(x, pmin, axis=axis, keepdims=True)
    ma = np.percentile(x, pmax, axis=axis, keepdims=True)
    return normalize_mi_ma(x, mi, ma, clip=clip, eps=eps, dtype=dtype)


def normalize_mi_ma(x, mi, ma, clip=False, eps=1e-20, dtype=np.float32):
    if dtype is not None:
        x = x.astype(dtype, copy=False)
        mi = dtype(mi) if np.isscalar(mi) else mi.astype(dtype, copy=False)
        ma = dtype(ma) if np.isscalar(ma) else ma.astype(dtype, copy=False)
        eps = dtype(eps)
    x = (x - mi) / (ma - mi + eps)
    if clip:
        x = np.clip(x, 0, 1)
    return x


def _is_floatarray(x):
    return isinstance(x.dtype.type(0), np.floating)
