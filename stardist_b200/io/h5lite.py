"""Minimal pure-Python HDF5 reader for Keras weight files (`weights_best.h5`, `model.save_weights`).

The reference loads its networks with `keras_model.load_weights(<basedir>/<name>/weights_best.h5)`
(csbdeep BaseModel._find_and_load_weights; stardist/models/base.py:210-228 builds the Keras graph whose layer
names are the group names in the file).  No HDF5 library is available on this path, and the files are simple:
h5py's default "earliest" layout -- superblock version 0, old-style groups (v1 B-tree + local heap + symbol table
nodes), version-1 object headers, contiguous (or compact) little-endian IEEE datasets without filters.  This reader
implements exactly that subset (HDF5 File Format Specification 2.0, sections II-IV) and raises on anything else
(chunked / filtered datasets, new-style groups, big-endian or exotic datatypes).
"""
import struct
import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"


class H5Error(ValueError):
    pass


class _File:
    def __init__(self, path):
        with open(path, "rb") as f:
            self.b = f.read()
        if self.b[:8] != _SIG:
            raise H5Error("not an HDF5 file: %s" % path)
        ver = self.b[8]
        if ver not in (0, 1):
            raise H5Error("unsupported superblock version %d (only the 'earliest' h5py layout is handled)" % ver)
        self.O = self.b[13]          # size of offsets
        self.L = self.b[14]          # size of lengths
        p = 24 if ver == 0 else 28   # after leaf/internal K, consistency flags (+ indexed-storage K, reserved in v1)
        self.base = self._off(p)
        p += 4 * self.O              # base address, free-space info, end of file, driver info
        # root group symbol table entry
        self.root = self._ste(p)

    def _int(self, p, n):
        return int.from_bytes(self.b[p:p + n], "little")

    def _off(self, p):
        return self._int(p, self.O)

    def _len(self, p):
        return self._int(p, self.L)

    def _ste(self, p):
        """symbol table entry -> dict(name_off, header, cache, btree, heap)"""
        O = self.O
        e = dict(name_off=self._off(p), header=self._off(p + O), cache=self._int(p + 2 * O, 4))
        if e["cache"] == 1:
            e["btree"] = self._off(p + 2 * O + 8)
            e["heap"] = self._off(p + 2 * O + 8 + O)
        return e

    # ------------------------------------------------------------------ object headers (version 1)
    def messages(self, addr):
        a = self.base + addr
        if self.b[a] != 1:
            raise H5Error("unsupported object header version %d" % self.b[a])
        n_msg = self._int(a + 2, 2)
        size = self._int(a + 8, 4)
        blocks = [(a + 16, size)]
        out = []
        while blocks and len(out) < n_msg:
            p, remaining = blocks.pop(0)
            end = p + remaining
            while p + 8 <= end and len(out) < n_msg:
                mtype, msize, flags = self._int(p, 2), self._int(p + 2, 2), self.b[p + 4]
                body = p + 8
                if mtype == 0x0010:      # continuation
                    blocks.append((self.base + self._off(body), self._len(body + self.O)))
                out.append((mtype, body, msize, flags))
                p = body + msize
                p = (p + 7) & ~7 if (p - (a + 16)) % 8 else p
        return out

    # ------------------------------------------------------------------ old-style groups
    def _heap_name(self, heap_addr, off):
        h = self.base + heap_addr
        if self.b[h:h + 4] != b"HEAP":
            raise H5Error("bad local heap signature")
        data = self.base + self._off(h + 8 + 2 * self.L)
        end = self.b.index(b"\x00", data + off)
        return self.b[data + off:end].decode("utf-8")

    def _btree_entries(self, addr, heap):
        t = self.base + addr
        if self.b[t:t + 4] != b"TREE":
            raise H5Error("bad B-tree signature")
        if self.b[t + 4] != 0:
            raise H5Error("not a group B-tree")
        level, used = self.b[t + 5], self._int(t + 6, 2)
        p = t + 8 + 2 * self.O
        p += self.L                       # key 0
        for _ in range(used):
            child = self._off(p)
            p += self.O + self.L
            if level > 0:
                yield from self._btree_entries(child, heap)
            else:
                s = self.base + child
                if self.b[s:s + 4] != b"SNOD":
                    raise H5Error("bad symbol table node signature")
                n = self._int(s + 6, 2)
                q = s + 8
                for _ in range(n):
                    e = self._ste(q)
                    yield self._heap_name(heap, e["name_off"]), e
                    q += 2 * self.O + 8 + 16

    def children(self, ste):
        """(name, symbol table entry) of a group; [] for a dataset"""
        if ste.get("cache") == 1:
            return list(self._btree_entries(ste["btree"], ste["heap"]))
        for mtype, body, msize, flags in self.messages(ste["header"]):
            if mtype == 0x0011:
                return list(self._btree_entries(self._off(body), self._off(body + self.O)))
            if mtype in (0x0002, 0x0006):
                raise H5Error("new-style groups (link messages) are not supported")
        return []

    # ------------------------------------------------------------------ datasets
    def dataset(self, header):
        shape = dtype = data = None
        for mtype, body, msize, flags in self.messages(header):
            if mtype == 0x0001:
                ver, rank = self.b[body], self.b[body + 1]
                p = body + (8 if ver == 1 else 4)
                shape = tuple(self._len(p + i * self.L) for i in range(rank))
            elif mtype == 0x0003:
                cls, ver = self.b[body] & 15, self.b[body] >> 4
                bits0 = self.b[body + 1]
                size = self._int(body + 4, 4)
                if bits0 & 1:
                    raise H5Error("big-endian datasets are not supported")
                if cls == 1 and size in (2, 4, 8):
                    dtype = np.dtype("<f%d" % size)
                elif cls == 0 and size in (1, 2, 4, 8):
                    dtype = np.dtype("<%s%d" % ("i" if (bits0 & 8) else "u", size))
                else:
                    raise H5Error("unsupported datatype class %d size %d" % (cls, size))
            elif mtype == 0x0008:
                ver = self.b[body]
                if ver == 3:
                    lclass = self.b[body + 1]
                    if lclass == 1:
                        data = ("contiguous", self._off(body + 2), self._len(body + 2 + self.O))
                    elif lclass == 0:
                        n = self._int(body + 2, 2)
                        data = ("compact", body + 4, n)
                    else:
                        raise H5Error("chunked datasets are not supported (Keras weight files are contiguous)")
                elif ver in (1, 2):
                    rank, lclass = self.b[body + 1], self.b[body + 2]
                    if lclass != 1:
                        raise H5Error("only contiguous datasets are supported")
                    data = ("contiguous", self._off(body + 8), None)
                else:
                    raise H5Error("unsupported data layout version %d" % ver)
            elif mtype == 0x000B:
                raise H5Error("filtered (compressed) datasets are not supported")
        if shape is None or dtype is None or data is None:
            return None
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        if data[0] == "contiguous":
            if data[1] == (1 << (8 * self.O)) - 1:
                return np.zeros(shape, dtype)          # never written
            start = self.base + data[1]
        else:
            start = data[1]
        return np.frombuffer(self.b, dtype=dtype, count=n, offset=start).reshape(shape).copy()


def read_datasets(path):
    """{'/group/.../name': ndarray} for every dataset of the file"""
    f = _File(path)
    out = {}

    def walk(prefix, ste, depth):
        if depth > 16:
            raise H5Error("group nesting too deep")
        kids = f.children(ste)
        if not kids:
            arr = f.dataset(ste["header"])
            if arr is not None:
                out[prefix] = arr
            return
        for name, e in kids:
            walk(prefix + "/" + name, e, depth + 1)

    walk("", f.root, 0)
    return out


def read_keras_weights(path):
    """Keras `save_weights` / `save` file -> {layer_name: (kernel, bias)} in the layout of stardist_b200's weights dict
    (kernel (k..., Cin, Cout), bias (Cout,)).  Datasets are `<layer>/<layer>/kernel:0` and `.../bias:0`, optionally
    below `/model_weights`."""
    ds = read_datasets(path)
    layers = {}
    for k, v in ds.items():
        parts = [p for p in k.split("/") if p]
        if parts and parts[0] == "model_weights":
            parts = parts[1:]
        if len(parts) < 2:
            continue
        layer, leaf = parts[0], parts[-1].split(":")[0]
        if leaf in ("kernel", "bias"):
            layers.setdefault(layer, {})[leaf] = np.ascontiguousarray(v, dtype=np.float32)
        elif leaf in ("gamma", "beta", "moving_mean", "moving_variance"):
            raise H5Error("batch-normalisation weights found (layer %s): not supported on this path" % layer)
    out = {}
    for layer, d in layers.items():
        if "kernel" not in d:
            continue
        k = d["kernel"]
        out[layer] = (k, d.get("bias", np.zeros(k.shape[-1], np.float32)))
    return out
