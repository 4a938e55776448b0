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

    def chunk_table(self):
        """(addr int64[n], nbytes int64[n], filter_mask uint32[n]) of the chunks in row order, for the native
        chunk streamer (``DeviceMatrix.append_h5_chunks``); None when the dataset is not row-chunked."""
        if self._layout["class"] != 2:
            return None
        chunk = self._layout["chunk"]
        if any(c != s for c, s in zip(chunk[1:], self.shape[1:])):
            return None
        index = self._chunk_index()
        if any(any(o != 0 for o in offs[1:]) for offs, _a, _n, _m in index):
            return None
        if [offs[0] for offs, _a, _n, _m in index] != list(range(0, len(index) * chunk[0], chunk[0])):
            return None                                              # holes / unordered chunks: use iter_chunks
        if self._filters and self._filters != [LZF_FILTER]:
            raise H5FormatError(f"unsupported filter pipeline {self._filters}")
        return (np.array([a for _o, a, _n, _m in index], dtype=np.int64),
                np.array([n for _o, _a, n, _m in index], dtype=np.int64),
                np.array([m for _o, _a, _n, m in index], dtype=np.uint32))

    @property
    def has_lzf(self):
        return bool(self._filters)

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
# writer: the same dialect (superblock v0, v1 headers, symbol-table root, chunked + LZF, v1 chunk B-trees)
# ------------------------------------------------------------------------------------------------
_BOOL_TYPE = bytes.fromhex("180200000100000010080000010000000000080046414c5345000000545255450000000000010000")
_F32_TYPE = bytes.fromhex("11201f000400000000002000170800177f00000000000000")
_I64_TYPE = bytes.fromhex("10080000080000000000400000000000")
_FILL = bytes.fromhex("0203000100000000")
_GROUP_LEAF_K, _GROUP_INTERNAL_K, _CHUNK_K = 4, 16, 32


def _pad8(b):
    return b + b"\x00" * (-len(b) % 8)


def _message(mtype, payload, flags=0):
    payload = _pad8(payload)
    return struct.pack("<HHB3x", mtype, len(payload), flags) + payload


def _auto_chunk_rows(n, typesize):
    """Chunk length h5py picks for a 1-D dataset created with ``chunks=True`` / a compression filter and no explicit
    chunk shape (its published ``guess_chunk`` heuristic: aim at 16 KB * 2^log10(bytes / 1 MB), clamped to 8 KB .. 1 MB,
    halving the length until the chunk is within 50 % of that) -- so `samples` and `var_count` get the chunk shape the
    reference's files have (utmos/select.py:209, :238; 2,504 samples -> chunks of 1,252)."""
    base, lo, hi = 16 * 1024, 8 * 1024, 1024 * 1024
    c = float(max(1, n))
    target = base * (2 ** np.log10(c * typesize / (1024.0 * 1024.0)))
    target = min(max(target, lo), hi)
    while True:
        nbytes = c * typesize
        if (nbytes < target or abs(nbytes - target) / target < 0.5) and nbytes < hi:
            break
        if c == 1:
            break
        c = np.ceil(c / 2.0)
    return int(c)


class _ChunkedDataset:
    """Bookkeeping of one chunked dataset while the file is being written."""

    def __init__(self, writer, dtype_msg, itemsize, row_shape, chunk_rows, unlimited):
        self.w = writer
        self.dtype_msg, self.itemsize = dtype_msg, itemsize
        self.row_shape = tuple(row_shape)            # trailing dims (empty for 1-D)
        self.chunk_rows = int(chunk_rows)
        self.unlimited = unlimited
        self.rows = 0
        self.chunks = []                             # (first_row, address, nbytes, filter_mask)
        self.pending = None

    @property
    def chunk_shape(self):
        return (self.chunk_rows,) + self.row_shape

    @property
    def chunk_nbytes(self):
        return int(np.prod(self.chunk_shape)) * self.itemsize

    def append(self, block):
        """block: ndarray [n, *row_shape] of the dataset dtype."""
        if block.shape[0] == 0:
            return
        if self.pending is not None and self.pending.shape[0]:
            block = np.concatenate([self.pending, block])
        n_full = block.shape[0] // self.chunk_rows
        for i in range(n_full):
            self._emit(block[i * self.chunk_rows:(i + 1) * self.chunk_rows])
        self.pending = np.ascontiguousarray(block[n_full * self.chunk_rows:])

    def _emit(self, rows):
        raw = np.ascontiguousarray(rows).tobytes()
        comp = _native.lzf_compress(raw)
        mask = 0
        if comp is None:                              # did not shrink: stored raw, filter skipped (bit 0)
            comp, mask = raw, 1
        addr = self.w._write(comp)
        self.chunks.append((self.rows, addr, len(comp), mask))
        self.rows += rows.shape[0]

    def append_encoded(self, blobs, rows):
        """Chunks that were encoded elsewhere (native encoder): [(bytes, filter_mask)] holding `rows` dataset rows."""
        if self.pending is not None and self.pending.shape[0]:
            raise H5FormatError("encoded chunks cannot follow a partial dense chunk")
        for k, (blob, mask) in enumerate(blobs):
            addr = self.w._write(blob)
            self.chunks.append((self.rows + k * self.chunk_rows, addr, len(blob), mask))
        self.rows += rows

    def finish(self):
        if self.pending is not None and self.pending.shape[0]:
            tail = self.pending
            full = np.zeros(self.chunk_shape, dtype=tail.dtype)
            full[:tail.shape[0]] = tail
            n = tail.shape[0]
            self._emit(full)
            self.rows += n - self.chunk_rows          # the padding rows are not part of the dataset
        self.pending = None

    # -- v1 B-tree of chunks ------------------------------------------------------------------
    def _key(self, nbytes, mask, first_row):
        offs = (first_row,) + (0,) * len(self.row_shape) + (0,)
        return struct.pack(f"<II{len(offs)}Q", nbytes, mask, *offs)

    def _end_key(self, first_row):
        offs = (first_row,) + self.row_shape + (self.itemsize,)
        return struct.pack(f"<II{len(offs)}Q", 0, 0, *offs)

    def write_btree(self):
        """Returns the root node address (UNDEF when there are no chunks)."""
        if not self.chunks:
            return UNDEF
        rank1 = len(self.row_shape) + 2
        key_size = 8 + 8 * rank1
        node_size = 24 + 2 * _CHUNK_K * (key_size + 8) + key_size
        cap = 2 * _CHUNK_K
        # level 0: (first key bytes, end row, child address) per chunk
        entries = [(self._key(nb, mask, first), first + self.chunk_rows, addr) for first, addr, nb, mask in self.chunks]
        level = 0
        while True:
            groups = [entries[i:i + cap] for i in range(0, len(entries), cap)]
            addrs = [self.w._reserve(node_size) for _ in groups]
            nxt = []
            for gi, grp in enumerate(groups):
                left = addrs[gi - 1] if gi > 0 else UNDEF
                right = addrs[gi + 1] if gi + 1 < len(groups) else UNDEF
                body = b"TREE" + struct.pack("<BBHQQ", 1, level, len(grp), left, right)
                for key, _end, child in grp:
                    body += key + struct.pack("<Q", child)
                body += self._end_key(grp[-1][1])
                self.w._write_at(addrs[gi], body.ljust(node_size, b"\x00"))
                nxt.append((grp[0][0], grp[-1][1], addrs[gi]))
            if len(nxt) == 1:
                return nxt[0][2]
            entries = nxt
            level += 1

    def header_messages(self, btree_addr):
        rank = 1 + len(self.row_shape)
        dims = (self.rows,) + self.row_shape
        maxd = ((UNDEF if self.unlimited else self.rows),) + self.row_shape
        space = struct.pack("<BBB5x", 1, rank, 1) + struct.pack(f"<{rank}Q", *dims) + struct.pack(f"<{rank}Q", *maxd)
        filt = (struct.pack("<BB6x", 1, 1) + struct.pack("<HHHH", LZF_FILTER, 8, 1, 3) + b"lzf".ljust(8, b"\x00") +
                struct.pack("<III", 4, 261, self.chunk_nbytes) + b"\x00" * 4)
        cdims = self.chunk_shape + (self.itemsize,)
        layout = struct.pack("<BBB", 3, 2, rank + 1) + struct.pack("<Q", btree_addr) + struct.pack(f"<{rank + 1}I", *cdims)
        return [_message(0x01, space), _message(0x03, self.dtype_msg, 1), _message(0x05, _FILL, 1),
                _message(0x0B, filt, 1), _message(0x08, layout)]


class H5Writer:
    """Creates the ``--lowmem NEW.hdf5`` file of utmos/select.py:198-238: datasets ``samples`` (fixed strings),
    ``data`` (bool, or float32 ``GT*AF``; chunks ``(max(1, int(1e6/4/S)), S)``, LZF) and ``var_count`` (int64)."""

    def __init__(self, path, samples, float_data=False):
        self.path = path
        self.samples = np.asarray(samples).astype("S")
        self.n_samples = len(self.samples)
        self.float_data = bool(float_data)
        self._fh = open(path, "wb")
        self._pos = 0
        self._pend_gt = self._pend_af = None                       # packed rows that do not fill a chunk yet
        self._write(b"\x00" * 96)                                  # superblock + root entry, patched in close()
        c_rows = max(1, int(1e6 / 4 / self.n_samples))             # utmos/select.py:205
        if self.float_data:
            self.data = _ChunkedDataset(self, _F32_TYPE, 4, (self.n_samples,), c_rows, True)
        else:
            self.data = _ChunkedDataset(self, _BOOL_TYPE, 1, (self.n_samples,), c_rows, True)

    # -- raw file access ------------------------------------------------------------------------
    def _write(self, payload):
        addr = self._pos
        self._fh.seek(addr)
        self._fh.write(payload)
        self._pos += len(payload)
        return addr

    def _reserve(self, nbytes):
        self._pos = (self._pos + 7) // 8 * 8
        addr = self._pos
        self._pos += nbytes
        return addr

    def _write_at(self, addr, payload):
        self._fh.seek(addr)
        self._fh.write(payload)

    # -- rows -------------------------------------------------------------------------------------
    def append_packed(self, gt_packed, af, threads=0):
        """One .jl part: informative rows only (utmos/select.py:275-280), dense bool or float32 GT*AF (:219-223).
        Whole chunks are unpacked and LZF-compressed by native host threads (csrc/hostio.cu); rows that do not fill
        a chunk wait, still packed, for the next part."""
        gt_packed = np.ascontiguousarray(gt_packed, dtype=np.uint8)
        af = None if af is None else np.asarray(af, dtype=np.float64).reshape(-1)
        if self.data.pending is not None and self.data.pending.shape[0]:
            return self.append_packed_py(gt_packed, af)          # dense rows are pending: stay on the dense path
        keep = gt_packed.any(axis=1)                             # pad bits are zero (np.packbits)
        gt_packed = gt_packed[keep]
        af_kept = af[keep] if self.float_data else None
        if self._pend_gt is not None and self._pend_gt.shape[0]:
            gt_packed = np.concatenate([self._pend_gt, gt_packed])
            if self.float_data:
                af_kept = np.concatenate([self._pend_af, af_kept])
        c_rows = self.data.chunk_rows
        batch = max(c_rows, (256 << 20) // max(1, self.n_samples * self.data.itemsize) // c_rows * c_rows)
        n_full = gt_packed.shape[0] // c_rows * c_rows
        for r0 in range(0, n_full, batch):
            r1 = min(n_full, r0 + batch)
            blobs = _native.h5_encode_chunks(gt_packed[r0:r1], self.n_samples, af_kept[r0:r1] if self.float_data else None,
                                             c_rows, threads)
            self.data.append_encoded(blobs, r1 - r0)
        self._pend_gt = gt_packed[n_full:].copy()
        self._pend_af = af_kept[n_full:].copy() if self.float_data else None
        return None

    def _flush_packed(self):
        """Rows that did not fill a chunk: one last, zero padded chunk (or back to the dense path)."""
        if self._pend_gt is None or not self._pend_gt.shape[0]:
            return
        gt, af = self._pend_gt, self._pend_af
        self._pend_gt = self._pend_af = None
        blobs = _native.h5_encode_chunks(gt, self.n_samples, af if self.float_data else None, self.data.chunk_rows)
        self.data.append_encoded(blobs, gt.shape[0])

    def append_packed_py(self, gt_packed, af):
        """NumPy restatement of append_packed (and the path taken when dense rows are pending)."""
        gt_packed = np.asarray(gt_packed)
        af = None if af is None else np.asarray(af, dtype=np.float64).reshape(-1)
        step = 8192
        for r0 in range(0, gt_packed.shape[0], step):
            dense = np.unpackbits(gt_packed[r0:r0 + step], axis=1, count=self.n_samples).astype(bool)
            keep = dense.any(axis=1)
            dense = dense[keep]
            if self.float_data:
                dense = (dense * af[r0:r0 + step][keep].reshape(-1, 1)).astype(np.float32)
            self.data.append(dense)

    def append_dense(self, block):
        if self._pend_gt is not None and self._pend_gt.shape[0]:
            # packed rows are waiting for their chunk to fill: continue densely from them
            gt, af = self._pend_gt, self._pend_af
            self._pend_gt = self._pend_af = None
            self.append_packed_py(gt, af if self.float_data else None)
        self.data.append(np.asarray(block, dtype=np.float32 if self.float_data else bool))

    # -- finish -----------------------------------------------------------------------------------
    def _write_header(self, messages):
        body = b"".join(messages)
        head = struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body))
        addr = self._reserve(16 + len(body))
        self._write_at(addr, head + body)
        return addr

    def close(self, var_count):
        self._flush_packed()
        self.data.finish()
        width = max(1, self.samples.dtype.itemsize)
        str_type = struct.pack("<BBBBI", 0x13, 0x01, 0, 0, width)
        ds_samples = _ChunkedDataset(self, str_type, width, (), _auto_chunk_rows(self.n_samples, width), True)
        ds_samples.append(self.samples.astype(f"S{width}"))
        ds_samples.finish()
        ds_vc = _ChunkedDataset(self, _I64_TYPE, 8, (), _auto_chunk_rows(self.n_samples, 8), False)
        ds_vc.append(np.asarray(var_count, dtype="<i8"))
        ds_vc.finish()
        self._pos = (self._pos + 7) // 8 * 8
        headers = {}
        for name, ds in (("data", self.data), ("samples", ds_samples), ("var_count", ds_vc)):
            headers[name] = self._write_header(ds.header_messages(ds.write_btree()))
        # local heap: names at 8-byte aligned offsets, offset 0 is the empty string
        names = sorted(headers)
        heap = bytearray(8)
        offsets = {}
        for name in names:
            offsets[name] = len(heap)
            heap += _pad8(name.encode() + b"\x00")
        free_off = len(heap)
        heap += struct.pack("<QQ", 1, 32) + b"\x00" * 16          # one free block of 32 bytes ends the segment
        heap_data = self._reserve(len(heap))
        self._write_at(heap_data, bytes(heap))
        heap_addr = self._reserve(32)
        self._write_at(heap_addr, b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), free_off, heap_data))
        # symbol table node + group B-tree
        snod_size = 8 + 2 * _GROUP_LEAF_K * 40
        snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
        for name in names:
            snod += struct.pack("<QQII16x", offsets[name], headers[name], 0, 0)
        snod_addr = self._reserve(snod_size)
        self._write_at(snod_addr, snod.ljust(snod_size, b"\x00"))
        tree_size = 24 + 2 * _GROUP_INTERNAL_K * 16 + 8
        tree = (b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) +
                struct.pack("<QQQ", 0, snod_addr, offsets[names[-1]]))
        tree_addr = self._reserve(tree_size)
        self._write_at(tree_addr, tree.ljust(tree_size, b"\x00"))
        root = self._write_header([_message(0x11, struct.pack("<QQ", tree_addr, heap_addr))])
        eof = self._pos
        sb = (SIGNATURE + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) +
              struct.pack("<HHI", _GROUP_LEAF_K, _GROUP_INTERNAL_K, 0) +
              struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF) +
              struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", tree_addr, heap_addr))
        assert len(sb) == 96
        self._write_at(0, sb)
        self._fh.truncate(eof)
        self._fh.close()
