"""CPU tests of the big-image block logic (stardist_b200/big.py) and of the multi-rank assembly
(stardist_b200/parallel_big.py, world_size 2 over gloo).  Mirrors the reference's tests/test_big.py:
test_cover2D/3D (:52-76, reassembling a label image from blocks is the identity), test_edgecases (:79)."""
import os, sys
import numpy as np
import pytest
from scipy import ndimage as ndi

from stardist_b200.big import Block, BlockND, NotFullyVisible
from stardist_b200.matching import relabel_sequential


def _label_image(shape, seed, rmax=6):
    rng = np.random.default_rng(seed)
    lbl = np.zeros(shape, np.int32)
    k = 0
    for _ in range(400):
        c = [rng.integers(rmax + 1, s - rmax - 1) for s in shape]
        r = rng.integers(2, rmax)
        sl = tuple(slice(ci - r, ci + r + 1) for ci in c)
        grids = np.ogrid[tuple(slice(-r, r + 1) for _ in shape)]
        ball = sum(g * g for g in grids) <= r * r
        if (lbl[sl][ball] > 0).any(): continue
        k += 1; lbl[sl][ball] = k
    return lbl


def _process_factory(gt):
    """fake predict_instances: the 'prediction' of a block is the GT labels inside its read region,
    objects cut by the read region's border removed, relabelled 1..n with a matching polys dict"""
    def process(block):
        axes = block.axes
        lab = block.read(gt, axes=axes).copy()
        lab = relabel_sequential(lab)[0].astype(np.int32)
        n = int(lab.max())
        objs = ndi.find_objects(lab)
        pts = np.array([[(s.start + s.stop) // 2 for s in sl] for sl in objs]).reshape(n, lab.ndim)
        polys = dict(prob=np.linspace(1, .5, n).astype(np.float32), points=pts)
        lab = block.crop_context(lab, axes=axes)
        return block.filter_objects(lab, polys, axes=axes)
    return process


def _serial(blocks, process, shape, axes):
    out = np.zeros(shape, np.int32); off = 1; pts = []
    for b in blocks:
        lab, polys = process(b)
        lab = relabel_sequential(lab, off)[0] if lab.max() > 0 else lab
        b.write(out, lab, axes=axes); pts.append(polys['points']); off += len(polys['prob'])
    return out, np.concatenate(pts)


@pytest.mark.parametrize("shape,axes,bs,mo,ctx,grid", [((160, 200), 'YX', 64, 16, 8, 1), ((150, 131), 'YX', 72, 18, 6, 3),
                                                        ((48, 90, 70), 'ZYX', (32, 48, 40), (12, 16, 14), (2, 4, 2), (1, 2, 2))])
def test_cover_reassembly_is_identity(shape, axes, bs, mo, ctx, grid):
    gt = _label_image(shape, seed=1)
    blocks = BlockND.cover(shape, axes, bs, mo, ctx, grid)
    out, pts = _serial(blocks, _process_factory(gt), shape, axes)
    assert np.array_equal(out > 0, gt > 0)
    # same partition into objects: one-to-one map between label ids
    pairs = np.unique(np.stack([gt[gt > 0], out[gt > 0]], 1), axis=0)
    assert len(pairs) == gt.max() == len(np.unique(out)) - 1 == len(pts)


def test_cover_edge_sizes():
    for size in range(7800, 7810):
        blocks = Block.cover(size, 1024, 128, 32, grid=8, verbose=False)
        assert blocks[0].start == 0 and blocks[-1].end == size
        w = [b.slice_write for b in blocks]
        assert w[0].start == 0 and w[-1].stop == size and all(a.stop >= b.start for a, b in zip(w[:-1], w[1:]))


def test_object_larger_than_overlap_raises():
    gt = np.zeros((64, 200), np.int32); gt[20:40, 10:190] = 1
    blocks = BlockND.cover(gt.shape, 'YX', (64, 80), (0, 16), (0, 4), 1)
    with pytest.raises(RuntimeError):
        _serial(blocks, _process_factory(gt), gt.shape, 'YX')


def _worker(rank, world, port, shape, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from stardist_b200 import parallel_big
    gt = _label_image(shape, seed=2)
    blocks = BlockND.cover(shape, 'YX', 64, 16, 8, 1)
    out, polys = parallel_big.run_sharded(blocks, _process_factory(gt), shape, 'YX', None, np.int32)
    if rank == 0:
        q.put((out, polys['points'], polys['prob']))
    dist.barrier(); dist.destroy_process_group()


@pytest.mark.parametrize("world,shape", [(2, (170, 210)), (8, (170, 210)), (8, (70, 100))])
def test_sharded_assembly_two_ranks_equals_serial(world, shape):
    """world 2 and 8 (the box size), and more ranks than blocks ((70, 100) has 6 blocks for 8 ranks: two idle ranks)"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, shape, q)) for r in range(world)]
    for p in procs: p.start()
    out, pts, prob = q.get(timeout=120)
    for p in procs: p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    gt = _label_image(shape, seed=2)
    blocks = BlockND.cover(shape, 'YX', 64, 16, 8, 1)
    want, want_pts = _serial(blocks, _process_factory(gt), shape, 'YX')
    assert np.array_equal(out, want) and np.array_equal(pts, want_pts) and len(prob) == len(pts)


def test_block_covers_equal_the_reference(golden_dir):
    """Block.cover / BlockND.cover vs tests/golden/big_cover.npz (the reference's stardist/big.py run by make_big_cover.py on
    the parameter sets of its own tests, tests/test_big.py:50-83, incl. the 7800..8000 edge sizes, and on configs[3] / [4]):
    identical read / crop / write slices for every block, same block order"""
    import json
    from stardist_b200.big import Block
    g = np.load(os.path.join(golden_dir, "big_cover.npz"))
    BIG = 10 ** 9
    for i, (size, bs, mo, ctx, grid) in enumerate(json.loads(bytes(g["cases_1d"]).decode())):
        blocks = Block.cover(size, bs, mo, ctx, grid, verbose=False)
        got = np.array([[b.start, b.end, b.slice_read.start, b.slice_read.stop, b.slice_crop_context.start,
                         b.slice_crop_context.stop if b.slice_crop_context.stop is not None else BIG,
                         b.slice_write.start, b.slice_write.stop, int(b.at_begin), int(b.at_end)] for b in blocks], np.int64)
        assert np.array_equal(got, g["1d/%d" % i]), (size, bs, mo, ctx, grid)
    for i, (shape, axes, bs, mo, ctx, grid) in enumerate(json.loads(bytes(g["cases_nd"]).decode())):
        tup = lambda v: tuple(v) if isinstance(v, list) else v
        blocks = BlockND.cover(tuple(shape), axes, tup(bs), tup(mo), tup(ctx), tup(grid))
        rows = []
        for b in blocks:
            row = [b.id]
            for sl in (b.slice_read(), b.slice_crop_context(), b.slice_write()):
                for s in sl: row += [s.start if s.start is not None else 0, s.stop if s.stop is not None else BIG]
            rows.append(row)
        assert np.array_equal(np.array(rows, np.int64), g["nd/%d" % i]), (shape, axes, bs, mo, ctx, grid)


def test_block_responsibility_equals_the_reference(golden_dir):
    """Block.is_responsible (big.py:89-122) on a lattice of query intervals per block vs the reference's decisions
    (tests/golden/big_cover.npz 'resp/*': 1 responsible, 0 not, 2 / 3 = NotFullyVisible(False / True))"""
    import json
    from stardist_b200.big import Block, NotFullyVisible
    g = np.load(os.path.join(golden_dir, "big_cover.npz"))
    n = 0
    for i, (size, bs, mo, ctx, grid) in enumerate(json.loads(bytes(g["cases_1d"]).decode())):
        blocks = Block.cover(size, bs, mo, ctx, grid, verbose=False)
        for k, a, e, want in g["resp/%d" % i]:
            try: got = int(bool(blocks[k].is_responsible((int(a), int(e)))))
            except NotFullyVisible as ex: got = 3 if ex.args[0] else 2
            assert got == want, (size, bs, mo, ctx, grid, k, a, e, got, want)
            n += 1
    assert n > 50000


def test_block_pipeline_equals_the_reference(golden_dir):
    """BlockND.read / crop_context / filter_objects / translate_coordinates / write, block by block, against the reference's
    own big.py on three synthetic label images (tests/golden/big_blocks.npz, make_big_blocks.py): the same objects survive in
    every block, with the same translated points, and the re-assembled image equals the input"""
    sys.path.insert(0, golden_dir)
    import make_big_blocks as mk
    g = np.load(os.path.join(golden_dir, "big_blocks.npz"))
    for ci, (shape, axes, bs, mo, ctx, grid, n, rmax, seed) in enumerate(mk.CASES):
        lab = g["%d/label" % ci]
        blocks = BlockND.cover(shape, axes, bs, mo, ctx, grid)
        result = np.zeros_like(lab)
        for b in blocks:
            local, polys, ids = mk.block_inputs(b, lab, axes)
            kept, polys_out = b.filter_objects(b.crop_context(local, axes=axes), polys, axes=axes)
            assert np.array_equal(kept, g["%d/%d/kept" % (ci, b.id)])
            for k in ("points", "prob", "dist"):
                assert np.array_equal(polys_out[k], g["%d/%d/%s" % (ci, b.id, k)]), (ci, b.id, k)
            assert np.array_equal(polys_out["rays_faces"], np.arange(6))
            b.write(result, np.where(kept > 0, b.crop_context(b.read(lab, axes=axes), axes=axes), 0), axes=axes)
        assert np.array_equal(result, g["%d/reassembled" % ci]) and np.array_equal(result, lab)


def test_predict_instances_big_bookkeeping_equals_the_reference(golden_dir):
    """StarDistBase.predict_instances_big (block cover, per-block predict_instances, crop, responsibility filter, running label
    offset, ordered write, poly concatenation; base.py:838-983) against the REFERENCE method run on the same stand-in model
    (tests/golden/big_blocks.npz 'big/*', make_big_blocks.py: context=None, per-axis grids, 2-D and 3-D)"""
    sys.path.insert(0, golden_dir)
    import make_big_blocks as mk
    from stardist_b200.models.base import StarDistBase
    g = np.load(os.path.join(golden_dir, "big_blocks.npz"))
    for ci, kw, grid, overlap in mk.BIG_CASES:
        axes = mk.CASES[ci][1]
        labels_out, polys = mk.run_big(StarDistBase.predict_instances_big, g["%d/label" % ci], axes, kw, grid, overlap)
        assert labels_out.dtype == g["big/%d/labels" % ci].dtype and np.array_equal(labels_out, g["big/%d/labels" % ci])
        for k in ("points", "prob", "dist", "rays_faces"):
            assert np.array_equal(polys[k], g["big/%d/%s" % (ci, k)]), (ci, k)
        assert set(polys) == {"points", "prob", "dist", "rays_faces"}


def test_vectorised_responsibility_equals_scalar_rule():
    """BlockND.responsible_many (used by the device block pipeline) == all(t.is_responsible(...)) incl. the
    NotFullyVisible cases and the short-circuit order over the axes (big.py:89-122, :340-345)"""
    import numpy as np
    from stardist_b200.big import BlockND, NotFullyVisible
    rng = np.random.default_rng(0)
    n_checked = 0
    for shape, axes, bs, mo, ctx in [((700, 640), 'YX', 256, 48, 16), ((150, 120, 130), 'ZYX', (64, 48, 56), (16, 8, 16), (8, 8, 4))]:
        for b in BlockND.cover(shape, axes, bs, mo, ctx, 1):
            wshape = [s.stop - s.start for s in b.slice_crop_context()]
            n = 400
            lo = np.stack([rng.integers(0, w, n) for w in wshape], 1)
            lo[rng.random(lo.shape) < 0.3] = 0
            hi = np.stack([np.minimum(w, lo[:, k] + 1 + rng.integers(0, 40, n)) for k, w in enumerate(wshape)], 1)
            hi = np.where(rng.random(hi.shape) < 0.25, np.array(wshape)[None], hi)
            mine, inv = b.responsible_many(lo, hi)
            for i in range(n):
                sl = tuple(slice(int(a), int(c)) for a, c in zip(lo[i], hi[i]))
                try:
                    r, e = b.is_responsible(sl), False
                except NotFullyVisible:
                    r, e = False, True
                assert e == inv[i] and (e or r == mine[i])
                n_checked += 1
    assert n_checked > 10000


def _worker_gather(rank, world, port, q):
    import torch, torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from stardist_b200 import parallel_big
    rng = np.random.default_rng(100 + rank)
    mine = {}
    for bid in range(rank, 7, world):                    # rank 2 of 3 owns blocks 2, 5; block counts incl. an empty block
        n = 0 if bid == 4 else int(rng.integers(1, 30))
        mine[bid] = dict(prob=rng.random(n).astype(np.float32), points=rng.integers(0, 1000, (n, 2)), coord=rng.random((n, 2, 32)).astype(np.float32),
                         extra="not-an-array-%d" % rank)
    got = parallel_big._gather_polys(mine, rank, world, torch.device("cpu"), None)
    mine_all = [None] * world if rank == 0 else None
    dist.gather_object(mine, mine_all, dst=0)
    if rank == 0:
        q.put((got, mine_all))
    dist.barrier(); dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_polygon_gather_as_raw_bytes_equals_object_gather(world):
    """parallel_big._gather_polys (object arrays as raw bytes point to point, only the description pickled) returns on rank 0
    exactly the per-block dicts a plain gather_object returns -- incl. empty blocks and ranks without blocks (world 8, 7 blocks)"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world
    procs = [ctx.Process(target=_worker_gather, args=(r, world, port, q)) for r in range(world)]
    for p in procs: p.start()
    got, mine_all = q.get(timeout=120)
    for p in procs: p.join(timeout=60)
    want = {}
    for d in mine_all: want.update(d)
    assert sorted(got) == sorted(want) == list(range(7))
    for b in want:
        assert set(got[b]) == set(want[b])
        for k in ('prob', 'points', 'coord'):
            assert got[b][k].dtype == want[b][k].dtype and np.array_equal(got[b][k], want[b][k]), (b, k)
