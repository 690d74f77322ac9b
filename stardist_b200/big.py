"""Block tiler for very large images (host-side integer bookkeeping).

Same contract as stardist/big.py -- `Block.cover` (:171-280) / `BlockND.cover` (:427-450) produce a
grid-aligned cover of overlapping blocks with a read region, a write region (read minus context) and a
responsibility rule (`is_responsible`, :89-122) such that every object smaller than `min_overlap` is
owned by exactly one block; `BlockND.read / crop_context / filter_objects / translate_coordinates /
write` (:312-425) are the per-block steps of predict_instances_big (models/base.py:953-972).

The implementation is array based (a cover is computed as integer vectors of starts / strides /
extra contexts and frozen into immutable records) rather than the reference's linked chain of mutable
blocks.  skimage.regionprops (un-vendored) is replaced by scipy.ndimage.find_objects, which yields the
same bounding boxes and object masks.  Multi-GPU sharding of the blocks: stardist_b200/parallel_big.py.
"""
import math
from itertools import product
import numpy as np
from scipy import ndimage as ndi
from .utils import axes_check_and_normalize

OBJECT_KEYS = set(('prob', 'points', 'coord', 'dist', 'class_prob', 'class_id'))
COORD_KEYS = set(('points', 'coord'))


class NotFullyVisible(Exception):
    pass


def _grid_divisible(grid, size, name=None, verbose=True):
    if size % grid == 0:
        return size
    up = int(math.ceil(size / grid) * grid)
    if bool(verbose):
        prefix = verbose if isinstance(verbose, str) else ''
        print(f"{prefix}increasing '{'value' if name is None else name}' from {size} to {up} to be evenly divisible by {grid} (grid)", flush=True)
    return up


class Block:
    """One axis-interval of a cover.  Immutable; all quantities in pixels.

    start/size: read region; context_start/context_end: discarded margins (0 at the image border);
    resp_start: first position (relative to the write region) this block is responsible for."""
    __slots__ = ('start', 'size', 'min_overlap', 'context', 'context_start', 'context_end', 'stride',
                 'at_begin', 'at_end', 'resp_start')

    def __init__(self, start, size, min_overlap, context, context_start, context_end, stride, at_begin, at_end, resp_start):
        self.start, self.size, self.min_overlap, self.context = int(start), int(size), int(min_overlap), int(context)
        self.context_start, self.context_end, self.stride = int(context_start), int(context_end), int(stride)
        self.at_begin, self.at_end, self.resp_start = bool(at_begin), bool(at_end), int(resp_start)

    @property
    def end(self):
        return self.start + self.size

    @property
    def overlap(self):
        return self.size - self.stride

    @property
    def slice_read(self):
        return slice(self.start, self.end)

    @property
    def slice_crop_context(self):
        return slice(self.context_start, self.size - self.context_end)

    @property
    def slice_write(self):
        return slice(self.start + self.context_start, self.end - self.context_end)

    def is_responsible(self, bbox):
        """bbox = (min, max) of an object relative to the block without context.  Exactly one block of a
        chain answers True for an object smaller than min_overlap; NotFullyVisible(spans_block) otherwise."""
        bmin, bmax = bbox
        r_end = self.size - self.context_start - self.context_end
        assert 0 <= bmin < bmax <= r_end
        if bmin == 0 and bmax >= self.resp_start:
            if bmax == r_end:
                raise NotFullyVisible(True)       # spans the whole block
            if not self.at_begin:
                raise NotFullyVisible(False)      # spans the whole overlap with the predecessor
        if bmax < self.resp_start:
            return False
        if bmax == r_end and not self.at_end:
            return False
        return True

    def responsible_many(self, bmin, bmax):
        """vectorised is_responsible for arrays of boxes (bmin inclusive, bmax exclusive, relative to the block without
        context): returns (mine bool[n], invisible bool[n]); invisible marks the boxes for which is_responsible raises"""
        bmin, bmax = np.asarray(bmin), np.asarray(bmax)
        r_end = self.size - self.context_start - self.context_end
        span = (bmin == 0) & (bmax >= self.resp_start)
        invisible = span & ((bmax == r_end) | (not self.at_begin))
        mine = ~(bmax < self.resp_start)
        if not self.at_end:
            mine &= ~(bmax == r_end)
        return mine, invisible

    def __repr__(self):
        w = self.slice_write
        return (f'Block({self.start:03}:{self.end:03}, write={w.start:03}:{w.stop:03}, '
                f'size={self.context_start}+{self.size-self.context_start-self.context_end}+{self.context_end})')

    @staticmethod
    def cover(size, block_size, min_overlap, context, grid=1, verbose=True):
        """Chain (list) of grid-aligned blocks covering [0, size]; only the last block may be shorter."""
        assert 0 <= min_overlap + 2 * context < block_size <= size
        assert 0 < grid <= block_size
        block_size = _grid_divisible(grid, block_size, name='block_size', verbose=verbose)
        min_overlap = _grid_divisible(grid, min_overlap, name='min_overlap', verbose=verbose)
        context = _grid_divisible(grid, context, name='context', verbose=verbose)
        size_orig = size
        size = _grid_divisible(grid, size, name='size', verbose=False)
        # ---- work in units of `grid`
        S, B, O, C = size // grid, block_size // grid, min_overlap // grid, context // grid
        assert 0 <= O + 2 * C < B       # (can be violated by the rounding to grid multiples above)
        stride0 = B - (O + 2 * C)
        n = 1
        while (n - 1) * stride0 + B < S:
            n += 1
        strides = [stride0] * n
        # shrink strides round-robin over the first n-1 blocks until the chain ends exactly at S
        excess = (n - 1) * stride0 + B - S
        i = 0
        while excess > 0:
            assert strides[i] > 1
            strides[i] -= 1
            excess -= 1
            i += 1
            if i == n - 1: i = 0
        starts = [0] * n
        for k in range(1, n):
            starts[k] = starts[k - 1] + strides[k - 1]
        extra_s, extra_e = [0] * n, [0] * n

        def ctx_s(k): return 0 if k == 0 else C + extra_s[k]
        def ctx_e(k): return 0 if k == n - 1 else C + extra_e[k]
        # write regions of non-neighbouring blocks must not overlap: split any excess between them
        for k in range(n - 2):
            ow = (starts[k] + B - ctx_e(k)) - (starts[k + 2] + ctx_s(k + 2))
            if ow > 0:
                extra_e[k] += ow // 2
                extra_s[k + 2] += ow - ow // 2
        # ---- back to pixels
        g = grid
        size_delta = size - size_orig
        assert 0 <= size_delta < grid
        blocks = []
        for k in range(n):
            sz = B * g - (size_delta if k == n - 1 else 0)
            cs, ce = ctx_s(k) * g, ctx_e(k) * g
            if k == 0:
                resp = 0
            else:
                pred_overlap = B * g - strides[k - 1] * g
                resp = pred_overlap - ctx_e(k - 1) * g - cs
            blocks.append(Block(starts[k] * g, sz, O * g, C * g, cs, ce, strides[k] * g, k == 0, k == n - 1, resp))
        # sanity checks (same invariants as the reference asserts)
        assert blocks[0].start == 0 and blocks[-1].end == size_orig
        for a, b in zip(blocks[:-1], blocks[1:]):
            assert a.overlap - 2 * C * g >= O * g
            assert a.slice_write.stop - b.slice_write.start >= O * g
            assert a.start % grid == 0 and a.end % grid == 0
        for a, c in zip(blocks[:-2], blocks[2:]):
            assert a.slice_write.stop <= c.slice_write.start
        return blocks


class BlockND:
    """N-dimensional block = one Block per axis + a unique id (Cartesian product of the 1-D covers)."""

    def __init__(self, id, blocks, axes):
        self.id = id
        self.blocks = tuple(blocks)
        self.axes = axes_check_and_normalize(axes, length=len(self.blocks))
        self.axis_to_block = dict(zip(self.axes, self.blocks))

    def blocks_for_axes(self, axes=None):
        axes = self.axes if axes is None else axes_check_and_normalize(axes)
        return tuple(self.axis_to_block[a] for a in axes)

    def slice_read(self, axes=None):
        return tuple(t.slice_read for t in self.blocks_for_axes(axes))

    def slice_crop_context(self, axes=None):
        return tuple(t.slice_crop_context for t in self.blocks_for_axes(axes))

    def slice_write(self, axes=None):
        return tuple(t.slice_write for t in self.blocks_for_axes(axes))

    def read(self, x, axes=None):
        return x[self.slice_read(axes)]

    def crop_context(self, labels, axes=None):
        return labels[self.slice_crop_context(axes)]

    def write(self, x, labels, axes=None):
        """write the entries > 0 of labels into the block's write region of x (later blocks overwrite)"""
        s = self.slice_write(axes)
        region = x[s]
        fg = labels > 0
        region[fg] = labels[fg]
        x[s] = region

    def is_responsible(self, slices, axes=None):
        return all(t.is_responsible((s.start, s.stop)) for t, s in zip(self.blocks_for_axes(axes), slices))

    def responsible_many(self, bmin, bmax, axes=None):
        """is_responsible for n boxes at once: bmin / bmax int arrays [n, ndim] (max exclusive).  Returns
        (mine bool[n], invisible bool[n]) with the evaluation order of `all(t.is_responsible(...) for t in blocks)`:
        an axis is only looked at while all previous axes answered True, so `invisible` (the NotFullyVisible cases of
        big.py:89-122) is raised by the first axis that objects before any axis has answered False."""
        blocks = self.blocks_for_axes(axes)
        bmin, bmax = np.asarray(bmin), np.asarray(bmax)
        n = len(bmin)
        alive = np.ones(n, bool)          # no axis has answered False (or raised) yet
        invisible = np.zeros(n, bool)
        for k, t in enumerate(blocks):
            m, inv = t.responsible_many(bmin[:, k], bmax[:, k])
            invisible |= alive & inv
            alive &= m & ~inv
        return alive, invisible

    def __repr__(self):
        return 'BlockND(%s|%s)' % (self.id, ','.join(f'{a}={t.start:03}:{t.end:03}' for t, a in zip(self.blocks, self.axes)))

    def __iter__(self):
        return iter(self.blocks)

    def filter_objects(self, labels, polys, axes=None):
        """Keep only the objects this block is responsible for.  `labels` is the context-cropped label
        image of the block, label id i <-> entry i-1 of `polys`.  Returns modified copies
        (labels_filtered, polys_out) -- coordinates translated to global positions -- or only the
        labels when polys is None.  RuntimeError if an object violates the min_overlap assumption."""
        assert np.issubdtype(labels.dtype, np.integer)
        blocks = self.blocks_for_axes(axes)
        ndim = len(blocks)
        assert ndim in (2, 3)
        assert labels.ndim == ndim and labels.shape == tuple(s.stop - s.start for s in self.slice_crop_context(axes))
        kept = np.zeros_like(labels)
        for lab, sl in enumerate(ndi.find_objects(labels), 1):
            if sl is None:
                continue
            try:
                mine = self.is_responsible(sl, axes)
            except NotFullyVisible:
                shape_object = tuple(s.stop - s.start for s in sl)
                shape_min_overlap = tuple(t.min_overlap for t in blocks)
                raise RuntimeError(f"Found object of shape {shape_object}, which violates the assumption of being smaller than 'min_overlap' {shape_min_overlap}. Increase 'min_overlap' to avoid this problem.")
            if mine:
                kept[sl][labels[sl] == lab] = lab
        if polys is None:
            return kept
        assert isinstance(polys, dict) and any(k in polys for k in COORD_KEYS)
        ids = np.unique(kept)
        ind = [i - 1 for i in ids if i > 0]
        out = {k: (v[ind] if k in OBJECT_KEYS else v) for k, v in polys.items()}
        for k in COORD_KEYS:
            if k in out:
                out[k] = self.translate_coordinates(out[k], axes=axes)
        return kept, out

    def translate_coordinates(self, coordinates, axes=None):
        """block-local coordinates (of the read region) -> global coordinates"""
        ndim = len(self.blocks_for_axes(axes))
        assert isinstance(coordinates, np.ndarray) and coordinates.ndim >= 2 and coordinates.shape[1] == ndim
        origin = np.array([s.start for s in self.slice_read(axes)])
        return coordinates + origin.reshape(tuple(ndim if d == 1 else 1 for d in range(coordinates.ndim)))

    @staticmethod
    def cover(shape, axes, block_size, min_overlap, context, grid=1):
        shape = tuple(shape)
        n = len(shape)
        axes = axes_check_and_normalize(axes, length=n)
        def _vec(v): return n * [v] if np.isscalar(v) else list(v)
        block_size, min_overlap, context, grid = _vec(block_size), _vec(min_overlap), _vec(context), _vec(grid)
        assert n == len(block_size) == len(min_overlap) == len(context) == len(grid)
        per_axis = [Block.cover(*args) for args in zip(shape, block_size, min_overlap, context, grid)]
        return tuple(BlockND(i, blocks, axes) for i, blocks in enumerate(product(*per_axis)))
