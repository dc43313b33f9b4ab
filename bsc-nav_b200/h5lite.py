"""Minimal HDF5 reader / writer for the reference's token file `feat.h5df` (memory_2.py:326-358, 185-200, 312-323).

The reference keeps its tokens in HDF5 through h5py (`h5py==3.12.1`, requirements.txt:58): one group `grid_{row}_{col}_{h}` per
voxel holding two extendible float32 datasets, `features (M, D)` and `distances (M,)`.  h5py / libhdf5 are not part of this
image, so the interchange is implemented here against the published HDF5 file format (version-0/1 superblock, version-1 object
headers, symbol-table groups = version-1 B-tree + local heap + SNOD leaves, data layout version 3: compact / contiguous /
chunked with a version-1 chunk B-tree) -- exactly what h5py writes with its default `libver='earliest'`.

Reader: any file of that family with fixed-point / floating-point little- or big-endian datasets, optional deflate (+shuffle)
filters, object-header continuation blocks, a user block in front of the superblock.  "New style" files (superblock 2/3 with
link messages / fractal heaps, `libver='latest'`) are rejected with a clear error.
Writer: whole-file writer (groups of datasets, one level) producing the same structures libhdf5 reads: extendible datasets
(`maxshape`) are chunked with one B-tree leaf, fixed ones contiguous.

API mirrors the slice of h5py the memory uses: File(path, 'r'|'w'), `in`, keys(), [name], create_group, create_dataset(name,
data=, maxshape=, chunks=), Dataset.shape / dtype / [...].
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
_LEAF_K, _NODE_K, _CHUNK_K = 4, 16, 32


class H5Error(ValueError):
    pass


# ----------------------------------------------------------------------------------------------------------- reader
class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        base = 0
        while True:                                   # superblock at 0, 512, 1024, ... (user block in front)
            if buf[base:base + 8] == SIG:
                break
            base = 512 if base == 0 else base * 2
            if base + 8 > len(buf):
                raise H5Error("not an HDF5 file (no superblock signature)")
        self.base = base
        ver = buf[base + 8]
        if ver not in (0, 1):
            raise H5Error(f"superblock version {ver}: only the 'earliest' file family (versions 0 / 1, symbol-table groups) is supported")
        so, sl = buf[base + 13], buf[base + 14]
        if (so, sl) != (8, 8):
            raise H5Error("only 8-byte offsets / lengths are supported")
        p = base + 24 + (4 if ver == 1 else 0)
        self.base_addr = self.u64(p)
        # root group symbol table entry
        ste = p + 32
        self.root_header = self.u64(ste + 8)

    def u16(self, o): return struct.unpack_from("<H", self.b, o)[0]
    def u32(self, o): return struct.unpack_from("<I", self.b, o)[0]
    def u64(self, o): return struct.unpack_from("<Q", self.b, o)[0]
    def at(self, addr): return self.base_addr + addr          # file addresses are relative to the superblock's base address

    # ---- object header (version 1) -> list of (type, flags, data offset, size)
    def messages(self, addr):
        o = self.at(addr)
        if self.b[o] != 1:
            raise H5Error(f"object header version {self.b[o]} (new-style file): unsupported")
        n_msg, size = self.u16(o + 2), self.u32(o + 8)
        out, blocks = [], [(o + 16, size)]
        while blocks and len(out) < n_msg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < n_msg:
                t, sz, fl = self.u16(p), self.u16(p + 2), self.b[p + 4]
                d = p + 8
                if t == 0x0010:                        # continuation: another block of messages
                    blocks.append((self.at(self.u64(d)), self.u64(d + 8)))
                out.append((t, fl, d, sz))
                p = d + sz
        return out

    # ---- groups
    def group_entries(self, header_addr):
        """{name: object header address} of an old-style group."""
        for t, _fl, d, _sz in self.messages(header_addr):
            if t == 0x0011:
                btree, heap = self.u64(d), self.u64(d + 8)
                return self._walk_group_btree(btree, self._heap_data(heap))
            if t in (0x0002, 0x0006):
                raise H5Error("new-style group (link messages): unsupported")
        raise H5Error("object is not a group")

    def _heap_data(self, addr):
        o = self.at(addr)
        if self.b[o:o + 4] != b"HEAP":
            raise H5Error("bad local heap signature")
        size, data = self.u64(o + 8), self.u64(o + 24)
        d = self.at(data)
        return self.b[d:d + size]

    def _walk_group_btree(self, addr, heap):
        out = {}
        if addr == UNDEF:
            return out
        o = self.at(addr)
        if self.b[o:o + 4] != b"TREE" or self.b[o + 4] != 0:
            raise H5Error("bad group B-tree node")
        level, n = self.b[o + 5], self.u16(o + 6)
        p = o + 24
        for i in range(n):
            child = self.u64(p + 8 + 16 * i)
            if level:
                out.update(self._walk_group_btree(child, heap))
            else:
                s = self.at(child)
                if self.b[s:s + 4] != b"SNOD":
                    raise H5Error("bad symbol table node")
                for j in range(self.u16(s + 6)):
                    e = s + 8 + 40 * j
                    no = self.u64(e)
                    name = heap[no:heap.index(b"\0", no)].decode()
                    out[name] = self.u64(e + 8)
        return out

    def is_group(self, header_addr):
        return any(t == 0x0011 for t, *_ in self.messages(header_addr))

    # ---- datasets
    def dataset_info(self, header_addr):
        shape = dtype = layout = None
        filters = []
        for t, _fl, d, sz in self.messages(header_addr):
            if t == 0x0001:
                ver, rank, flags = self.b[d], self.b[d + 1], self.b[d + 2]
                p = d + (8 if ver == 1 else 4)
                shape = tuple(self.u64(p + 8 * i) for i in range(rank))
                maxshape = tuple(self.u64(p + 8 * (rank + i)) for i in range(rank)) if flags & 1 else shape
            elif t == 0x0003:
                dtype = self._dtype(d)
            elif t == 0x0008:
                layout = self._layout(d)
            elif t == 0x000B:
                filters = self._filters(d)
        if shape is None or dtype is None or layout is None:
            raise H5Error("object is not a (supported) dataset")
        return shape, tuple(None if m == UNDEF else m for m in maxshape), dtype, layout, filters

    def _dtype(self, d):
        cls, ver = self.b[d] & 0x0F, self.b[d] >> 4
        bits0 = self.b[d + 1]
        size = self.u32(d + 4)
        order = ">" if bits0 & 1 else "<"
        if cls == 0:
            return np.dtype(f"{order}{'i' if bits0 & 8 else 'u'}{size}")
        if cls == 1:
            return np.dtype(f"{order}f{size}")
        raise H5Error(f"datatype class {cls} (version {ver}): only integers and floats are supported")

    def _layout(self, d):
        ver = self.b[d]
        if ver in (1, 2):                              # libhdf5 <= 1.6 files: version, dimensionality, class, 5 reserved, [address], dims
            nd, cls = self.b[d + 1], self.b[d + 2]
            p = d + 8
            addr = None
            if cls != 0:
                addr = self.u64(p)
                p += 8
            dims = tuple(self.u32(p + 4 * i) for i in range(nd))
            if cls == 0:
                n = self.u32(p + 4 * nd)
                return ("compact", p + 4 * nd + 4, n)
            if cls == 1:
                return ("contiguous", addr, None)
            return ("chunked", addr, dims)
        if ver != 3:
            raise H5Error(f"data layout version {ver}: unsupported")
        cls = self.b[d + 1]
        if cls == 0:
            n = self.u16(d + 2)
            return ("compact", d + 4, n)
        if cls == 1:
            return ("contiguous", self.u64(d + 2), self.u64(d + 10))
        if cls == 2:
            nd = self.b[d + 2]
            return ("chunked", self.u64(d + 3), tuple(self.u32(d + 11 + 4 * i) for i in range(nd)))
        raise H5Error("unknown layout class")

    def _filters(self, d):
        ver, n = self.b[d], self.b[d + 1]
        p = d + (8 if ver == 1 else 2)
        out = []
        for _ in range(n):
            fid = self.u16(p)
            if ver == 1 or fid >= 256:
                name_len = self.u16(p + 2)
                p += 2
            else:
                name_len = 0
            n_cd = self.u16(p + 4)
            p += 6 + name_len
            cd = [self.u32(p + 4 * i) for i in range(n_cd)]
            p += 4 * n_cd
            if ver == 1 and n_cd % 2:
                p += 4
            out.append((fid, cd))
        return out

    def read_dataset(self, header_addr) -> np.ndarray:
        shape, _max, dtype, layout, filters = self.dataset_info(header_addr)
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        if layout[0] == "compact":
            raw = self.b[layout[1]:layout[1] + layout[2]]
            return np.frombuffer(raw, dtype=dtype, count=n).reshape(shape).copy()
        if layout[0] == "contiguous":
            if layout[1] == UNDEF:
                return np.zeros(shape, dtype)
            o = self.at(layout[1])
            return np.frombuffer(self.b, dtype=dtype, count=n, offset=o).reshape(shape).copy()
        _cls, btree, cdims = layout
        chunk = cdims[:-1]
        out = np.zeros(shape, dtype)
        if btree != UNDEF:
            for offs, addr, nbytes, mask in self._walk_chunk_btree(btree, len(cdims)):
                raw = self.b[self.at(addr):self.at(addr) + nbytes]
                for i, (fid, cd) in reversed(list(enumerate(filters))):
                    if mask & (1 << i):
                        continue
                    if fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 2:
                        es = cd[0] if cd else dtype.itemsize
                        a = np.frombuffer(raw, np.uint8)
                        m = len(a) // es
                        raw = a[:m * es].reshape(es, m).T.tobytes() + a[m * es:].tobytes()
                    else:
                        raise H5Error(f"filter {fid}: unsupported")
                blk = np.frombuffer(raw, dtype=dtype, count=int(np.prod(chunk))).reshape(chunk)
                sel = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, shape))
                if any(s.start >= s.stop for s in sel):
                    continue
                out[sel] = blk[tuple(slice(0, s.stop - s.start) for s in sel)]
        return out

    def _walk_chunk_btree(self, addr, nd):
        o = self.at(addr)
        if self.b[o:o + 4] != b"TREE" or self.b[o + 4] != 1:
            raise H5Error("bad chunk B-tree node")
        level, n = self.b[o + 5], self.u16(o + 6)
        ksz = 8 + 8 * nd
        p = o + 24
        for i in range(n):
            k = p + i * (ksz + 8)
            child = self.u64(k + ksz)
            if level:
                yield from self._walk_chunk_btree(child, nd)
            else:
                offs = tuple(self.u64(k + 8 + 8 * j) for j in range(nd - 1))
                yield offs, child, self.u32(k), self.u32(k + 4)


class Dataset:
    def __init__(self, name, data=None, maxshape=None, reader=None, addr=None):
        self.name, self._data, self.maxshape, self._r, self._addr = name, data, maxshape, reader, addr
        if reader is not None:
            self._shape, self.maxshape, self.dtype, _l, _f = reader.dataset_info(addr)

    @property
    def shape(self):
        return self._shape if self._r is not None else self._data.shape

    def _load(self):
        if self._data is None:
            self._data = self._r.read_dataset(self._addr)
        return self._data

    def __getitem__(self, k):
        return self._load()[k]

    def __array__(self, dtype=None, copy=None):
        a = self._load()
        return a.astype(dtype) if dtype is not None else a

    def __len__(self):
        return self.shape[0]


class Group:
    def __init__(self, name="/", reader=None, addr=None):
        self.name, self._r, self._addr = name, reader, addr
        self._items = {}                 # writer: name -> Group | Dataset
        self._entries = None             # reader: name -> header address

    def _names(self):
        if self._r is None:
            return sorted(self._items)
        if self._entries is None:
            self._entries = self._r.group_entries(self._addr)
        return sorted(self._entries)                       # HDF5 iterates links of old-style groups in name order

    def keys(self):
        return self._names()

    def __iter__(self):
        return iter(self._names())

    def __len__(self):
        return len(self._names())

    def __contains__(self, name):
        return name in (self._items if self._r is None else (self._names() and self._entries))

    def __getitem__(self, name):
        if self._r is None:
            return self._items[name]
        self._names()
        if name not in self._items:
            addr = self._entries[name]
            self._items[name] = Group(name, self._r, addr) if self._r.is_group(addr) else Dataset(name, reader=self._r, addr=addr)
        return self._items[name]

    # ---- writer side
    def create_group(self, name):
        if self._r is not None:
            raise H5Error("file is open read-only")
        if name in self._items:
            raise ValueError(f"Unable to create group (name already exists): {name}")
        g = self._items[name] = Group(name)
        return g

    def create_dataset(self, name, data=None, maxshape=None, chunks=None, dtype=None, shape=None):
        if self._r is not None:
            raise H5Error("file is open read-only")
        if name in self._items:
            raise ValueError(f"Unable to create dataset (name already exists): {name}")
        a = np.asarray(data if data is not None else np.zeros(shape, dtype or np.float32))
        if dtype is not None:
            a = a.astype(dtype)
        if a.dtype.kind not in "fiu":
            raise H5Error("only integer / float datasets are supported")
        d = self._items[name] = Dataset(name, data=np.ascontiguousarray(a.astype(a.dtype.newbyteorder("<"))), maxshape=maxshape)
        return d


class File(Group):
    """File(path, 'r') reads; File(path, 'w') collects groups / datasets in memory and writes the file on close."""

    def __init__(self, path, mode="r"):
        if mode not in ("r", "w"):
            raise H5Error("h5lite supports modes 'r' and 'w' (whole-file writer)")
        self.path, self.mode = path, mode
        if mode == "r":
            with open(path, "rb") as f:
                r = _Reader(f.read())
            super().__init__("/", r, r.root_header)
        else:
            super().__init__("/")

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self):
        if self.mode == "w" and self._items is not None:
            data = _Writer().build(self)
            tmp = self.path + ".tmp"
            with open(tmp, "wb") as f:
                f.write(data)
            import os
            os.replace(tmp, self.path)                  # atomic: a reader never sees a half-written file
            self._items = None


# ----------------------------------------------------------------------------------------------------------- writer
def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


def _msg(mtype: int, data: bytes, flags: int = 0) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


class _Writer:
    def __init__(self):
        self.out = bytearray()

    def alloc(self, data: bytes) -> int:
        self.out += b"\0" * (-len(self.out) % 8)
        addr = len(self.out)
        self.out += data
        return addr

    def header(self, msgs: list) -> int:
        body = b"".join(msgs)
        return self.alloc(struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body)) + body)

    def dataset(self, ds: Dataset) -> int:
        a = ds._data
        rank = a.ndim
        ext = ds.maxshape is not None and any(m is None or m != s for m, s in zip(ds.maxshape, a.shape))
        space = struct.pack("<BBB5x", 1, rank, 1 if ds.maxshape is not None else 0) + b"".join(struct.pack("<Q", s) for s in a.shape)
        if ds.maxshape is not None:
            space += b"".join(struct.pack("<Q", UNDEF if m is None else m) for m in ds.maxshape)
        if a.dtype.kind == "f":
            if a.dtype.itemsize == 4:
                props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            elif a.dtype.itemsize == 8:
                props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            else:
                raise H5Error("float16 datasets are not supported")
            dt = struct.pack("<BBBBI", 0x11, 0x20, 8 * a.dtype.itemsize - 1, 0, a.dtype.itemsize) + props
        else:
            dt = struct.pack("<BBBBI", 0x10, 0x08 if a.dtype.kind == "i" else 0x00, 0, 0, a.dtype.itemsize) + struct.pack("<HH", 0, 8 * a.dtype.itemsize)
        raw = a.tobytes()
        if ext:
            # extendible: chunked, one chunk holding the whole current extent (chunk dims >= 1), one B-tree leaf
            cdims = tuple(max(int(s), 1) for s in a.shape)
            chunk = np.zeros(cdims, a.dtype)
            chunk[tuple(slice(0, s) for s in a.shape)] = a
            caddr = self.alloc(chunk.tobytes())
            ksz = 8 + 8 * (rank + 1)
            node = bytearray(struct.pack("<4sBBHQQ", b"TREE", 1, 0, 1, UNDEF, UNDEF))
            node += struct.pack("<II", chunk.nbytes, 0) + b"".join(struct.pack("<Q", 0) for _ in range(rank + 1))
            node += struct.pack("<Q", caddr)
            node += struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", c) for c in cdims) + struct.pack("<Q", 0)   # key past the last chunk
            node += b"\0" * (24 + (2 * _CHUNK_K + 1) * ksz + 2 * _CHUNK_K * 8 - len(node))
            baddr = self.alloc(bytes(node))
            layout = struct.pack("<BBB", 3, 2, rank + 1) + struct.pack("<Q", baddr) + b"".join(struct.pack("<I", c) for c in cdims) + struct.pack("<I", a.dtype.itemsize)
            fill = struct.pack("<BBBB", 2, 3, 2, 0)      # allocate incrementally, write fill value if set, none defined
        else:
            daddr = self.alloc(raw) if raw else UNDEF
            layout = struct.pack("<BB", 3, 1) + struct.pack("<QQ", daddr, len(raw))
            fill = struct.pack("<BBBB", 2, 2, 2, 0)      # allocate late
        return self.header([_msg(0x0001, space), _msg(0x0003, dt, flags=1), _msg(0x0005, fill), _msg(0x0008, layout)])

    def group(self, grp: Group) -> tuple:
        """Writes the group's children, its local heap, SNOD leaves and B-tree; returns (header address, btree, heap)."""
        names = sorted(grp._items, key=lambda s: s.encode())
        child_addr = {}
        for n in names:
            it = grp._items[n]
            child_addr[n] = self.group(it)[0] if isinstance(it, Group) else self.dataset(it)
        # local heap: "" at offset 0, then the names (8-byte aligned, null terminated)
        heap = bytearray(8)
        off = {}
        for n in names:
            off[n] = len(heap)
            heap += _pad8(n.encode() + b"\0")
        free = len(heap)
        heap += struct.pack("<QQ", 1, 16)                     # one free block at the end (next = 1: end of list; size 16)
        daddr = self.alloc(bytes(heap))
        haddr = self.alloc(struct.pack("<4sB3xQQQ", b"HEAP", 0, len(heap), free, daddr))
        # leaves: SNOD nodes of up to 2 * leaf K entries, sorted by name
        leaves = []
        per = 2 * _LEAF_K
        for i in range(0, len(names), per):
            part = names[i:i + per]
            node = bytearray(struct.pack("<4sBBH", b"SNOD", 1, 0, len(part)))
            for n in part:
                it = grp._items[n]
                if isinstance(it, Group):
                    node += struct.pack("<QQII", off[n], child_addr[n], 1, 0) + struct.pack("<QQ", it._bt, it._hp)
                else:
                    node += struct.pack("<QQII16x", off[n], child_addr[n], 0, 0)
            node += b"\0" * (8 + per * 40 - len(node))
            leaves.append((self.alloc(bytes(node)), off[part[-1]]))
        # B-tree levels: nodes of up to 2 * node K children; key i+1 = heap offset of the largest name under child i
        level, nodes = 0, leaves
        if not nodes:
            nodes = []
        while True:
            per_n = 2 * _NODE_K
            groups = [nodes[i:i + per_n] for i in range(0, max(len(nodes), 1), per_n)]
            built = []
            for gi, part in enumerate(groups):
                node = bytearray(struct.pack("<4sBBHQQ", b"TREE", 0, level, len(part), UNDEF, UNDEF))
                node += struct.pack("<Q", 0)
                for caddr, last in part:
                    node += struct.pack("<QQ", caddr, last)
                node += b"\0" * (24 + (2 * per_n + 1) * 8 - len(node))
                built.append([self.alloc(bytes(node)), part[-1][1] if part else 0])
            # sibling links
            for i, (addr, _l) in enumerate(built):
                left = built[i - 1][0] if i else UNDEF
                right = built[i + 1][0] if i + 1 < len(built) else UNDEF
                struct.pack_into("<QQ", self.out, addr + 8, left, right)
            if len(built) == 1:
                bt = built[0][0]
                break
            nodes, level = [tuple(b) for b in built], level + 1
        grp._bt, grp._hp = bt, haddr
        hdr = self.header([_msg(0x0011, struct.pack("<QQ", bt, haddr))])
        return hdr, bt, haddr

    def build(self, root: Group) -> bytes:
        self.out = bytearray(96)                                # superblock, filled in last
        hdr, bt, hp = self.group(root)
        eof = len(self.out) + (-len(self.out) % 8)
        self.out += b"\0" * (eof - len(self.out))
        sb = SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, _LEAF_K, _NODE_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, hdr, 1, 0) + struct.pack("<QQ", bt, hp)
        self.out[0:96] = sb
        return bytes(self.out)
