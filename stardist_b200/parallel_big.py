"""Multi-GPU predict_instances_big: shard the block cover over the ranks of one node.

Reference semantics (stardist/models/base.py:953-975): blocks are processed in id order, each block
contributes `n_b` objects, label ids are offset by the running sum of the previous blocks' counts
(relabel_sequential(labels, label_offset), :959,:972), label tiles are written in id order so that
later blocks win inside overlaps (big.py:319-326) and the polygon dicts are concatenated in id order.
The blocks themselves are independent, so:

  rank r processes blocks r, r+W, r+2W, ...      (one process per GPU, no data-path collective)
  all-reduce(SUM) of the per-block object counts   -> exclusive scan in id order = label offsets
  label tiles + polygon arrays are sent to rank 0 in block-id order (point-to-point; NCCL over NVLink
  for CUDA tensors, gloo for the CPU tests), rank 0 assembles exactly like the serial loop.

The exchanged volume is small (counts: 8 B/block; tiles: 4 B/pixel once; polygons: ~140 B each), so
this is latency- not bandwidth-bound; nothing here is a compute kernel.
"""
import numpy as np
import torch
import torch.distributed as dist
from .big import OBJECT_KEYS
from .matching import relabel_sequential


def rank_world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def _dev(group=None):
    backend = dist.get_backend(group)
    return torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")


def _send_array(a, dst, dev, group):
    a = np.ascontiguousarray(a)
    hdr = np.zeros(10, np.int64)
    hdr[0] = a.ndim; hdr[1:1 + a.ndim] = a.shape; hdr[9] = _DT.index(a.dtype.str)
    dist.send(torch.from_numpy(hdr).to(dev), dst, group=group)
    if a.size:
        dist.send(torch.from_numpy(a.view(np.uint8).reshape(-1)).to(dev), dst, group=group)


def _recv_array(src, dev, group):
    hdr = torch.zeros(10, dtype=torch.int64, device=dev)
    dist.recv(hdr, src, group=group)
    hdr = hdr.cpu().numpy()
    shape = tuple(int(v) for v in hdr[1:1 + int(hdr[0])])
    dt = np.dtype(_DT[int(hdr[9])])
    n = int(np.prod(shape)) * dt.itemsize
    if n == 0:
        return np.zeros(shape, dt)
    buf = torch.empty(n, dtype=torch.uint8, device=dev)
    dist.recv(buf, src, group=group)
    return buf.cpu().numpy().view(dt).reshape(shape)


_DT = ['<f4', '<f8', '<i4', '<i8', '|u1', '|b1', '<i2', '<u2', '<u4', '<u8', '<f2']


def run_sharded(blocks, process, shape_out, axes_out, labels_out, labels_out_dtype, group=None):
    """process(block) -> (labels_cropped_filtered int array, polys dict).  Returns (labels_out, polys_all)
    on rank 0 and (None, None) elsewhere."""
    rank, world = rank_world(group)
    dev = _dev(group)
    nb = len(blocks)
    mine = {}
    counts = torch.zeros(nb, dtype=torch.int64)
    for b in blocks:
        if b.id % world != rank:
            continue
        labels, polys = process(b)
        mine[b.id] = (labels, polys)
        counts[b.id] = len(polys['prob'])
    counts = counts.to(dev)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    counts = counts.cpu().numpy()
    offsets = 1 + np.concatenate([[0], np.cumsum(counts)[:-1]])
    # relabel locally with the global offsets
    for bid, (labels, polys) in list(mine.items()):
        mine[bid] = (relabel_sequential(labels, int(offsets[bid]))[0] if labels.size and labels.max() > 0 else labels, polys)
    want_labels = not (np.isscalar(labels_out) and bool(labels_out) is False)
    keys = None
    if rank == 0:
        if want_labels and labels_out is None:
            labels_out = np.zeros(shape_out, dtype=labels_out_dtype)
        polys_all = {}
    for b in blocks:
        owner = b.id % world
        if rank == 0:
            if owner == 0:
                labels, polys = mine[b.id]
            else:
                labels = _recv_array(owner, dev, group) if want_labels else None
                nk = torch.zeros(1, dtype=torch.int64, device=dev); dist.recv(nk, owner, group=group)
                polys = {}
                for _ in range(int(nk.item())):
                    kname = bytes(_recv_array(owner, dev, group).tolist()).decode()
                    polys[kname] = _recv_array(owner, dev, group)
            if want_labels:
                b.write(labels_out, labels.astype(labels_out.dtype, copy=False), axes=axes_out)
            for k, v in polys.items():
                polys_all.setdefault(k, []).append(v)
        elif owner == rank:
            labels, polys = mine[b.id]
            if want_labels:
                _send_array(labels, 0, dev, group)
            arrs = {k: np.asarray(v) for k, v in polys.items() if isinstance(v, np.ndarray)}
            dist.send(torch.tensor([len(arrs)], dtype=torch.int64, device=dev), 0, group=group)
            for k, v in arrs.items():
                _send_array(np.frombuffer(k.encode(), np.uint8), 0, dev, group)
                _send_array(v, 0, dev, group)
    if rank != 0:
        return None, None
    polys_all = {k: (np.concatenate(v) if k in OBJECT_KEYS else v[0]) for k, v in polys_all.items()}
    return (labels_out if want_labels else None), polys_all


def block_offsets(counts):
    """label offsets of base.py:959,972: 1 + exclusive scan of the per-block object counts in block-id order"""
    counts = np.asarray(counts, dtype=np.int64)
    return 1 + np.concatenate([[0], np.cumsum(counts)[:-1]]) if len(counts) else np.zeros(0, np.int64)


def run_sharded_device(blocks, process_device, shape_out, axes_out, want_labels=True, group=None):
    """Device-resident assembly of predict_instances_big.

    process_device(block) -> (tile int32 CUDA tensor of the block's write region with ids 1..n_kept, polys dict, n_kept).
    Rank r owns the blocks r, r+W, ... (no data-path collective while they are processed).  Then
      * all-reduce(SUM) of the per-block object counts -> exclusive scan = label offsets (base.py:959,972),
      * the owners' tiles travel device-to-device to rank 0 (batched NCCL send/recv over NVLink, straight out of and into
        HBM; no host staging), rank 0 scatters every tile into the device-resident global map in block-id order with the
        offset folded in (sdb_label_write; later blocks win inside overlaps, big.py:319-326),
      * the polygon dicts (a few hundred bytes per object) are gathered with gather_object and concatenated in id order.
    Returns (labels int32 CUDA tensor | None, polys_all) on rank 0, (None, None) elsewhere.  world == 1: same code
    path without the collectives."""
    import ctypes, os, time
    from . import _lib as L
    lib = L.require_cuda()
    rank, world = rank_world(group)
    dev = torch.device("cuda", torch.cuda.current_device())
    timing = os.environ.get("STARDIST_B200_BIG_TIMING") == "1"
    T = [time.perf_counter()]
    def lap(name):
        if timing:
            torch.cuda.synchronize(); T.append(time.perf_counter())
            print("[big rank %d] %-22s %.1f ms" % (rank, name, 1e3 * (T[-1] - T[-2])), flush=True)
    nb = len(blocks)
    nd = len(shape_out)
    mine = {}
    counts = torch.zeros(nb, dtype=torch.int64)
    glob = None
    if rank == 0 and want_labels:
        glob = torch.zeros(tuple(int(s) for s in shape_out), dtype=torch.int32, device=dev)
    offsets_known = world == 1
    running = 1
    for b in blocks:
        if b.id % world != rank:
            continue
        tile, polys, n_kept = process_device(b)
        counts[b.id] = n_kept
        if offsets_known:
            # single process: the running offset is known, write at once and drop the tile
            if want_labels:
                origin = [s.start for s in b.slice_write(axes_out)]
                L.check(lib.sdb_label_write(L.ptr(tile), nd, L.iarr(tile.shape), running - 1, L.ptr(glob), L.iarr(shape_out),
                                           L.iarr(origin), L.stream_ptr()))
            running += n_kept
            mine[b.id] = (None, polys)
        else:
            mine[b.id] = (tile if want_labels else None, polys)
    lap("blocks (%d)" % len(mine))
    if world > 1:
        counts_d = counts.to(dev)
        dist.all_reduce(counts_d, op=dist.ReduceOp.SUM, group=group)
        counts = counts_d.cpu()
        lap("all-reduce counts")
    offsets = block_offsets(counts.numpy())
    if world > 1 and want_labels:
        ops, remote = [], {}
        if rank == 0:
            for b in blocks:
                owner = b.id % world
                if owner != 0:
                    shp = tuple(s.stop - s.start for s in b.slice_write(axes_out))
                    remote[b.id] = torch.empty(shp, dtype=torch.int32, device=dev)
                    ops.append(dist.P2POp(dist.irecv, remote[b.id], owner, group=group))
        else:
            for bid in sorted(mine):
                ops.append(dist.P2POp(dist.isend, mine[bid][0], 0, group=group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        if rank == 0:
            for b in blocks:
                tile = mine[b.id][0] if b.id % world == 0 else remote[b.id]
                origin = [s.start for s in b.slice_write(axes_out)]
                L.check(lib.sdb_label_write(L.ptr(tile), nd, L.iarr(tile.shape), int(offsets[b.id]) - 1, L.ptr(glob), L.iarr(shape_out),
                                           L.iarr(origin), L.stream_ptr()))
    lap("tiles -> rank 0 + scatter")
    polys_by_block = {bid: p for bid, (_, p) in mine.items()}
    if world > 1:
        polys_by_block = _gather_polys(polys_by_block, rank, world, dev, group)
        lap("polygons -> rank 0")
        if rank != 0:
            return None, None
    polys_all = {}
    for b in blocks:
        for k, v in polys_by_block[b.id].items():
            polys_all.setdefault(k, []).append(v)
    polys_all = {k: (np.concatenate(v) if k in OBJECT_KEYS else v[0]) for k, v in polys_all.items()}
    return glob, polys_all


def _gather_polys(polys_by_block, rank, world, dev, group):
    """per-block polygon dicts of all ranks -> rank 0 ({block id: dict}; None elsewhere).  The object arrays (coord / dist,
    points, prob, ...: ~300 B per object, tens of MB for a large image) travel as raw bytes through NCCL point-to-point --
    one buffer per rank and key, straight from / into device memory; only the small description (dtypes, shapes, per-block
    counts, non-array entries) goes through gather_object (pickle)."""
    bids = sorted(polys_by_block)
    keys = sorted(k for k in OBJECT_KEYS if bids and k in polys_by_block[bids[0]] and isinstance(polys_by_block[bids[0]][k], np.ndarray))
    packed, meta = {}, dict(bids=bids, keys={}, other={})
    for k in keys:
        arrs = [np.ascontiguousarray(polys_by_block[b][k]) for b in bids]
        cat = np.concatenate(arrs) if arrs else np.zeros(0)
        meta['keys'][k] = (cat.dtype.str, tuple(cat.shape[1:]), [len(a) for a in arrs])
        packed[k] = torch.from_numpy(cat.view(np.uint8).reshape(-1)).to(dev) if cat.size else None
    if bids:
        meta['other'] = {k: v for k, v in polys_by_block[bids[0]].items() if k not in keys}
    metas = [None] * world if rank == 0 else None
    dist.gather_object(meta, metas, dst=0, group=group)
    ops, recv = [], {}
    if rank == 0:
        for r in range(1, world):
            for k, (dt, tail, counts) in metas[r]['keys'].items():
                nbytes = int(sum(counts)) * int(np.prod(tail, dtype=np.int64)) * np.dtype(dt).itemsize
                if nbytes:
                    recv[(r, k)] = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                    ops.append(dist.P2POp(dist.irecv, recv[(r, k)], r, group=group))
    else:
        for k in keys:
            if packed[k] is not None:
                ops.append(dist.P2POp(dist.isend, packed[k], 0, group=group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    if rank != 0:
        return None
    out = dict(polys_by_block)
    for r in range(1, world):
        m = metas[r]
        parts = {}
        for k, (dt, tail, counts) in m['keys'].items():
            buf = recv.get((r, k))
            flat = buf.cpu().numpy().view(np.dtype(dt)) if buf is not None else np.zeros(0, np.dtype(dt))
            arr = flat.reshape((-1,) + tuple(tail))
            offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
            parts[k] = [arr[offs[i]:offs[i + 1]] for i in range(len(counts))]
        for i, b in enumerate(m['bids']):
            d = {k: parts[k][i] for k in m['keys']}
            d.update(m['other'])
            out[b] = d
    return out
