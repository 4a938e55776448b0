"""Minimal reader / writer for the hdf5 dialect utmos writes with h5py 3.7.

h5py / libhdf5 are not part of this stack.  The files the reference creates (utmos/select.py:198-238) use a
tiny, fixed subset of the format -- superblock v0, v1 object headers, a symbol-table root group, chunked
datasets indexed by a v1 B-tree, one filter (id 32000, LZF) -- so a few hundred lines are enough to stream
their chunks to the GPU and to write files stock h5py can read back.

Reader: ``H5File(path)`` -> ``f['data']`` has ``.shape``, ``.dtype``, ``.chunks``, ``iter_chunks()`` (row blocks
in row order, decoded through the native LZF codec) and ``read()``.
Writer: ``H5Writer(path)`` appends row blocks to ``data`` and writes ``samples`` / ``var_count`` on close.
"""
import struct

import numpy as np

from utmos_b200 import _native

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
LZF_FILTER = 32000


class H5FormatError(ValueError):
    """The file uses a feature outside the supported dialect."""


# ------------------------------------------------------------------------------------------------
# reader
# ------------------------------------------------------------------------------------------------
class Dataset:
    """One hdf5 dataset (contiguous or chunked + optional LZF)."""

    def __init__(self, fh, name, shape, dtype, layout, filters):
        self._fh = fh
        self.name = name
        self.shape = tuple(shape)
        self.dtype = dtype
        self._layout = layout
        self._filters = filters
        self.chunks = layout.get("chunk") if layout["class"] == 2 else None

    def _chunk_index(self):
        """[(offsets tuple, address, nbytes, filter_mask)] sorted by offsets."""
        out = []
        self._walk_btree(self._layout["btree"], out)
        out.sort(key=lambda c: c[0])
        return out

    def _walk_btree(self, addr, out):
        if addr == UNDEF:
            return
        fh = self._fh
        fh.seek(addr)
        head = fh.read(24)
        if head[:4] != b"TREE":
            raise H5FormatError("chunk index is not a v1 B-tree")
        node_type, level, used = head[4], head[5], struct.unpack_from("<H", head, 6)[0]
        if node_type != 1:
            raise H5FormatError("expected a raw-data chunk B-tree")
        rank1 = len(self._layout["chunk"]) + 1
        key_size = 8 + 8 * rank1
        body = fh.read(used * (key_size + 8) + key_size)
        pos = 0
        for _ in range(used):
            nbytes, fmask = struct.unpack_from("<II", body, pos)
            offs = struct.unpack_from(f"<{rank1}Q", body, pos + 8)
            child = struct.unpack_from("<Q", body, pos + key_size)[0]
            pos += key_size + 8
            if level > 0:
                self._walk_btree(child, out)
            else:
                out.append((offs[:-1], child, nbytes, fmask))

    def _decode_chunk(self, addr, nbytes, fmask):
        self._fh.seek(addr)
        raw = self._fh.read(nbytes)
        chunk_bytes = int(np.prod(self._layout["chunk"])) * self.dtype.itemsize
        if self._filters and not fmask & 1:
            if self._filters != [LZF_FILTER]:
                raise H5FormatError(f"unsupported filter pipeline {self._filters}")
            return _native.lzf_decompress(raw, chunk_bytes)
        return np.frombuffer(raw, dtype=np.uint8, count=chunk_bytes)

    def iter_chunks(self):
        """Yield (first_row, ndarray[rows, ...]) in row order; rows beyond the dataset end are trimmed."""
        if self._layout["class"] == 1:                               # contiguous
            yield 0, self.read()
            return
        chunk = self._layout["chunk"]
        if any(c != s for c, s in zip(chunk[1:], self.shape[1:])):
            raise H5FormatError("chunks must span whole rows")
        for offs, addr, nbytes, fmask in self._chunk_index():
            if any(o != 0 for o in offs[1:]):
                raise H5FormatError("chunks must span whole rows")
            block = self._decode_chunk(addr, nbytes, fmask).view(self.dtype).reshape(chunk)
            rows = min(chunk[0], self.shape[0] - offs[0])
            if rows > 0:
                yield offs[0], block[:rows]

    def read(self):
        """Whole dataset as one ndarray."""
        if self._layout["class"] == 1:
            if self._layout["addr"] == UNDEF:
                return np.zeros(self.shape, dtype=self.dtype)
            self._fh.seek(self._layout["addr"])
            count = int(np.prod(self.shape))
            return np.frombuffer(self._fh.read(count * self.dtype.itemsize), dtype=self.dtype).reshape(self.shape)
        out = np.zeros(self.shape, dtype=self.dtype)
        for first, block in self.iter_chunks():
            out[first:first + block.shape[0]] = block
        return out

    def __getitem__(self, key):
        return self.read()[key]


class H5File:
    """Read-only view of an hdf5 file in the utmos dialect."""

    def __init__(self, path):
        self.path = path
        self._fh = open(path, "rb")
        self._datasets = {}
        self._parse()

    def close(self):
        self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __getitem__(self, name):
        return self._datasets[name]

    def __contains__(self, name):
        return name in self._datasets

    def keys(self):
        return self._datasets.keys()

    # -- structure walk ------------------------------------------------------------------------
    def _parse(self):
        fh = self._fh
        head = fh.read(56 + 40)
        if head[:8] != SIGNATURE:
            raise H5FormatError("not an hdf5 file")
        if head[8] != 0:
            raise H5FormatError(f"superblock version {head[8]} not supported (expected 0)")
        if head[13] != 8 or head[14] != 8:
            raise H5FormatError("only 8-byte offsets/lengths are supported")
        # root symbol table entry starts at byte 56: name offset, header addr, cache type, reserved, scratch
        root_header = struct.unpack_from("<Q", head, 56 + 8)[0]
        msgs = self._read_header(root_header)
        sym = [m for m in msgs if m[0] == 0x11]
        if not sym:
            raise H5FormatError("root group is not a symbol-table group")
        btree, heap = struct.unpack_from("<QQ", sym[0][1], 0)
        heap_data = self._read_heap(heap)
        for name, addr in self._walk_group(btree, heap_data):
            ds = self._read_dataset(name, addr)
            if ds is not None:
                self._datasets[name] = ds

    def _read_heap(self, addr):
        self._fh.seek(addr)
        head = self._fh.read(32)
        if head[:4] != b"HEAP":
            raise H5FormatError("bad local heap")
        size, _free, data_addr = struct.unpack_from("<QQQ", head, 8)
        self._fh.seek(data_addr)
        return self._fh.read(size)

    def _walk_group(self, addr, heap):
        fh = self._fh
        fh.seek(addr)
        head = fh.read(24)
        if head[:4] != b"TREE" or head[4] != 0:
            raise H5FormatError("bad group B-tree")
        level, used = head[5], struct.unpack_from("<H", head, 6)[0]
        body = fh.read(used * 16 + 8)
        children = [struct.unpack_from("<Q", body, 8 + 16 * i)[0] for i in range(used)]
        out = []
        for child in children:
            if level > 0:
                out.extend(self._walk_group(child, heap))
                continue
            fh.seek(child)
            snod = fh.read(8)
            if snod[:4] != b"SNOD":
                raise H5FormatError("bad symbol table node")
            count = struct.unpack_from("<H", snod, 6)[0]
            entries = fh.read(40 * count)
            for i in range(count):
                name_off, obj = struct.unpack_from("<QQ", entries, 40 * i)
                end = heap.index(b"\x00", name_off)
                out.append((heap[name_off:end].decode(), obj))
        return out

    def _read_header(self, addr):
        """v1 object header -> [(type, payload bytes, flags)] following continuation blocks."""
        fh = self._fh
        fh.seek(addr)
        head = fh.read(16)
        if head[0] != 1:
            raise H5FormatError(f"object header version {head[0]} not supported (expected 1)")
        nmsg = struct.unpack_from("<H", head, 2)[0]
        size = struct.unpack_from("<I", head, 8)[0]
        blocks = [(addr + 16, size)]
        msgs = []
        while blocks and len(msgs) < nmsg:
            baddr, bsize = blocks.pop(0)
            fh.seek(baddr)
            data = fh.read(bsize)
            pos = 0
            while pos + 8 <= len(data) and len(msgs) < nmsg:
                mtype, msize, mflags = struct.unpack_from("<HHB", data, pos)
                payload = data[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x10:
                    blocks.append(struct.unpack_from("<QQ", payload, 0))
                msgs.append((mtype, payload, mflags))
        return msgs

    def _read_dataset(self, name, addr):
        msgs = self._read_header(addr)
        shape = dtype = layout = None
        filters = []
        for mtype, payload, _flags in msgs:
            if mtype == 0x01:
                shape = _parse_dataspace(payload)
            elif mtype == 0x03:
                dtype = _parse_datatype(payload)
            elif mtype == 0x08:
                layout = _parse_layout(payload)
            elif mtype == 0x0B:
                filters = _parse_filters(payload)
        if shape is None or dtype is None or layout is None:
            return None
        return Dataset(self._fh, name, shape, dtype, layout, filters)


def _parse_dataspace(p):
    version, rank, _flags = p[0], p[1], p[2]
    if version == 1:
        base = 8
    elif version == 2:
        base = 4
    else:
        raise H5FormatError(f"dataspace version {version}")
    return struct.unpack_from(f"<{rank}Q", p, base)


def _parse_datatype(p):
    cls, version = p[0] & 0x0F, p[0] >> 4
    bits0 = p[1]
    size = struct.unpack_from("<I", p, 4)[0]
    if cls == 0:                                      # fixed point
        signed = bool(bits0 & 0x08)
        if bits0 & 0x01:
            raise H5FormatError("big-endian integers not supported")
        return np.dtype(f"<{'i' if signed else 'u'}{size}")
    if cls == 1:                                      # floating point
        if bits0 & 0x01:
            raise H5FormatError("big-endian floats not supported")
        return np.dtype(f"<f{size}")
    if cls == 3:                                      # fixed-length string
        return np.dtype(f"S{size}")
    if cls == 8:                                      # enum (h5py bool = enum over int8 FALSE/TRUE)
        if version not in (1, 2, 3) or size != 1:
            raise H5FormatError("only 1-byte enums (h5py bool) are supported")
        return np.dtype(bool)
    raise H5FormatError(f"datatype class {cls} not supported")


def _parse_layout(p):
    if p[0] != 3:
        raise H5FormatError(f"layout version {p[0]} not supported (expected 3)")
    cls = p[1]
    if cls == 1:
        addr, size = struct.unpack_from("<QQ", p, 2)
        return {"class": 1, "addr": addr, "size": size}
    if cls == 2:
        ndim = p[2]
        btree = struct.unpack_from("<Q", p, 3)[0]
        dims = struct.unpack_from(f"<{ndim}I", p, 11)
        return {"class": 2, "btree": btree, "chunk": tuple(dims[:-1]), "elem": dims[-1]}
    raise H5FormatError("compact layout not supported")


def _parse_filters(p):
    if p[0] != 1:
        raise H5FormatError(f"filter pipeline version {p[0]}")
    count = p[1]
    pos = 8
    out = []
    for _ in range(count):
        fid, name_len, _flags, nvals = struct.unpack_from("<HHHH", p, pos)
        pos += 8 + (name_len + 7) // 8 * 8 + 4 * nvals + (4 if nvals % 2 else 0)
        out.append(fid)
    return out


# ------------------------------------------------------------------------------------------------
# writer (placeholder until the dialect writer lands)
# ------------------------------------------------------------------------------------------------
class H5Writer:
    """Creates the ``--lowmem NEW.hdf5`` file (utmos/select.py:198-238)."""

    def __init__(self, path, samples, float_data=False):
        raise NotImplementedError("writing --lowmem hdf5 files is not implemented yet; load the inputs directly")
