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
