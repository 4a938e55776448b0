"""Strict structural reader for the hdf5 dialect of utmos's --lowmem files.  TEST INFRASTRUCTURE.

Independent of utmos_b200/h5lite.py on purpose: it is what the tests use to show that a file written by
``h5lite.H5Writer`` has, message by message, the structure of a file h5py 3.7 / libhdf5 wrote for the same content
(the reference's fixtures tiny.hdf5 / tiny.af.hdf5), and that every on-disk invariant of the format specification
(HDF5 File Format Specification v3, sections II.A superblock v0, III.A.1 v1 B-trees, III.B group symbol nodes, III.D local
heaps, IV.A.1 v1 object headers, IV.A.2 messages) holds.  Addresses are checked for consistency, never compared.
"""
import struct

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"


class Bad(AssertionError):
    pass


def need(cond, what):
    if not cond:
        raise Bad(what)


class File:
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.b = fh.read()
        self.size = len(self.b)
        self.superblock = self._superblock()
        self.datasets = self._group(self.superblock["root_header"])

    def at(self, addr, n):
        need(addr != UNDEF and addr + n <= self.size, f"read [{addr}, +{n}) outside the file ({self.size} bytes)")
        return self.b[addr:addr + n]

    # -- II.A superblock version 0 -------------------------------------------------------------------------
    def _superblock(self):
        b = self.at(0, 96)
        need(b[:8] == SIGNATURE, "signature")
        sb = {"version": b[8], "freespace_version": b[9], "root_group_version": b[10], "reserved0": b[11],
              "shared_header_version": b[12], "size_of_offsets": b[13], "size_of_lengths": b[14], "reserved1": b[15]}
        sb["group_leaf_k"], sb["group_internal_k"], sb["consistency_flags"] = struct.unpack_from("<HHI", b, 16)
        base, freespace, eof, driver = struct.unpack_from("<QQQQ", b, 24)
        sb.update(base_address=base, freespace_address=freespace, eof_address=eof, driver_address=driver)
        name_off, header, cache_type, _res = struct.unpack_from("<QQII", b, 56)
        sb.update(root_name_offset=name_off, root_header=header, root_cache_type=cache_type)
        sb["root_scratch"] = struct.unpack_from("<QQ", b, 80)
        need(sb["version"] == 0 and sb["size_of_offsets"] == 8 and sb["size_of_lengths"] == 8, "superblock v0 with 8-byte fields")
        need(eof == self.size, f"end-of-file address {eof} != file size {self.size}")
        need(base == 0 and freespace == UNDEF and driver == UNDEF, "base / free-space / driver addresses")
        return sb

    # -- IV.A.1 version 1 object header ----------------------------------------------------------------------
    def header(self, addr):
        b = self.at(addr, 16)
        need(b[0] == 1 and b[1] == 0, "object header version 1")
        nmsg, refcount, size = struct.unpack_from("<HII", b, 2)
        need(refcount >= 1, "object reference count")
        blocks = [(addr + 16, size)]
        msgs = []
        while blocks:
            baddr, bsize = blocks.pop(0)
            data = self.at(baddr, bsize)
            pos = 0
            while pos + 8 <= len(data):
                mtype, msize, mflags = struct.unpack_from("<HHB", data, pos)
                need(msize % 8 == 0, f"message 0x{mtype:x} size {msize} is not a multiple of 8")
                need(pos + 8 + msize <= len(data), f"message 0x{mtype:x} runs past its header block")
                payload = data[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x10:
                    blocks.append(struct.unpack_from("<QQ", payload, 0))
                msgs.append((mtype, mflags, payload))
        need(len(msgs) == nmsg, f"header says {nmsg} messages, found {len(msgs)}")
        return msgs

    # -- III.B groups: symbol table message -> v1 B-tree (type 0) + local heap + SNOD ----------------------
    def _group(self, header_addr):
        msgs = self.header(header_addr)
        sym = [m for m in msgs if m[0] == 0x11]
        need(len(sym) == 1, "root group: one symbol table message")
        btree, heap = struct.unpack_from("<QQ", sym[0][2], 0)
        need((btree, heap) == self.superblock["root_scratch"] or self.superblock["root_cache_type"] == 0,
             "root entry scratch pad = (B-tree, heap) when its cache type is 1")
        hb = self.at(heap, 32)
        need(hb[:4] == b"HEAP" and hb[4] == 0, "local heap signature / version")
        hsize, hfree, hdata = struct.unpack_from("<QQQ", hb, 8)
        heap_bytes = self.at(hdata, hsize)
        need(hfree == UNDEF or hfree + 16 <= hsize, "local heap free-list head inside the data segment")
        names = {}
        self._group_node(btree, heap_bytes, names)
        out = {}
        for name, addr in names.items():
            out[name] = self._dataset(addr)
        self.group_entries = sorted(names)
        return out

    def _group_node(self, addr, heap, names):
        b = self.at(addr, 24)
        need(b[:4] == b"TREE" and b[4] == 0, "group B-tree node signature / type 0")
        level, used = b[5], struct.unpack_from("<H", b, 6)[0]
        need(1 <= used <= 2 * self.superblock["group_internal_k"], "group B-tree entries used")
        body = self.at(addr + 24, used * 16 + 8)
        for i in range(used):
            child = struct.unpack_from("<Q", body, 8 + 16 * i)[0]
            if level > 0:
                self._group_node(child, heap, names)
                continue
            sn = self.at(child, 8)
            need(sn[:4] == b"SNOD" and sn[4] == 1, "symbol table node signature / version")
            count = struct.unpack_from("<H", sn, 6)[0]
            need(1 <= count <= 2 * self.superblock["group_leaf_k"], "symbols in a node")
            ent = self.at(child + 8, 40 * count)
            prev = None
            for k in range(count):
                name_off, obj, cache_type, _r = struct.unpack_from("<QQII", ent, 40 * k)
                end = heap.index(b"\x00", name_off)
                name = heap[name_off:end].decode()
                need(prev is None or prev < name, "symbols sorted by name")
                need(cache_type == 0, "dataset entries carry no cached metadata")
                prev = name
                names[name] = obj

    # -- IV.A.2 dataset messages ---------------------------------------------------------------------------------
    def _dataset(self, addr):
        d = {"messages": []}
        for mtype, mflags, p in self.header(addr):
            if mtype == 0x00:
                continue                                          # NIL: padding
            d["messages"].append((mtype, mflags))
            if mtype == 0x01:
                need(p[0] == 1, "dataspace version 1")
                rank, flags = p[1], p[2]
                dims = struct.unpack_from(f"<{rank}Q", p, 8)
                maxd = struct.unpack_from(f"<{rank}Q", p, 8 + 8 * rank) if flags & 1 else None
                d["dataspace"] = {"rank": rank, "flags": flags, "dims": dims, "maxdims": maxd}
            elif mtype == 0x03:
                size = struct.unpack_from("<I", p, 4)[0]
                d["datatype"] = {"class": p[0] & 15, "version": p[0] >> 4, "bits": tuple(p[1:4]), "size": size,
                                 "properties": bytes(p[8:]).rstrip(b"\x00")}
            elif mtype == 0x05:
                d["fill"] = {"version": p[0], "alloc_time": p[1], "write_time": p[2], "defined": p[3]}
            elif mtype == 0x0B:
                need(p[0] == 1, "filter pipeline version 1")
                pos, filters = 8, []
                for _ in range(p[1]):
                    fid, name_len, flags, nvals = struct.unpack_from("<HHHH", p, pos)
                    name = p[pos + 8:pos + 8 + name_len].rstrip(b"\x00").decode()
                    need(name_len % 8 == 0, "filter name padded to 8 bytes")
                    vals = struct.unpack_from(f"<{nvals}I", p, pos + 8 + name_len)
                    pos += 8 + name_len + 4 * nvals + (4 if nvals % 2 else 0)
                    filters.append({"id": fid, "name": name, "flags": flags, "client": vals})
                d["filters"] = filters
            elif mtype == 0x08:
                need(p[0] == 3, "layout version 3")
                d["layout"] = {"class": p[1]}
                if p[1] == 2:
                    ndim = p[2]
                    btree = struct.unpack_from("<Q", p, 3)[0]
                    d["layout"]["dims"] = struct.unpack_from(f"<{ndim}I", p, 11)
                    d["layout"]["btree"] = btree
                elif p[1] == 1:
                    d["layout"]["addr"], d["layout"]["size"] = struct.unpack_from("<QQ", p, 2)
        need({"dataspace", "datatype", "layout"} <= set(d), "dataset has dataspace, datatype and layout messages")
        if d["layout"]["class"] == 2:
            d["chunks"] = self._chunk_tree(d)
        return d

    # -- III.A.1 v1 B-tree of raw data chunks (node type 1) ------------------------------------------------------
    def _chunk_tree(self, d):
        cdims = d["layout"]["dims"]                               # chunk dims + element size
        rank1 = len(cdims)
        need(rank1 == d["dataspace"]["rank"] + 1 and cdims[-1] == d["datatype"]["size"], "chunk dims = rank + element size")
        key_size = 8 + 8 * rank1
        root = d["layout"]["btree"]
        chunks = []
        if root == UNDEF:
            return chunks
        levels = {}

        def node(addr, lo_key, hi_key):
            b = self.at(addr, 24)
            need(b[:4] == b"TREE" and b[4] == 1, "chunk B-tree node signature / type 1")
            level, used = b[5], struct.unpack_from("<H", b, 6)[0]
            left, right = struct.unpack_from("<QQ", b, 8)
            need(1 <= used <= 64, "chunk B-tree entries used (K = 32)")
            levels.setdefault(level, []).append((addr, left, right))
            body = self.at(addr + 24, used * (key_size + 8) + key_size)
            keys = []
            for i in range(used + 1):
                nbytes, fmask = struct.unpack_from("<II", body, i * (key_size + 8))
                offs = struct.unpack_from(f"<{rank1}Q", body, i * (key_size + 8) + 8)
                keys.append((nbytes, fmask, offs))
            for i in range(used):
                need(keys[i][2] < keys[i + 1][2], "chunk keys strictly ascending")
                need(keys[i][2][-1] == 0, "element offset of a chunk key is 0")
                need(all(o % c == 0 for o, c in zip(keys[i][2][:-1], cdims[:-1])), "chunk offsets are multiples of the chunk dims")
            if lo_key is not None:
                need(keys[0][2] == lo_key, "first key of a child = its key in the parent")
            if hi_key is not None:
                need(keys[used][2] <= hi_key, "last key of a child <= the next key in the parent")
            for i in range(used):
                child = struct.unpack_from("<Q", body, i * (key_size + 8) + key_size)[0]
                if level > 0:
                    node(child, keys[i][2], keys[i + 1][2])
                else:
                    nbytes, fmask, offs = keys[i]
                    need(nbytes > 0 and child + nbytes <= self.size, "chunk lies inside the file")
                    chunks.append({"offset": offs[:-1], "nbytes": nbytes, "filter_mask": fmask, "addr": child})

        node(root, None, None)
        for level, nodes in levels.items():                        # sibling pointers chain every level left to right
            for i, (addr, left, right) in enumerate(nodes):
                need(left == (nodes[i - 1][0] if i else UNDEF), f"left sibling at level {level}")
                need(right == (nodes[i + 1][0] if i + 1 < len(nodes) else UNDEF), f"right sibling at level {level}")
        need([c["offset"] for c in chunks] == sorted(c["offset"] for c in chunks), "chunks in key order")
        return chunks


def comparable(f):
    """What must be EQUAL between two files of the same content (everything but addresses and compressed sizes)."""
    sb = {k: v for k, v in f.superblock.items() if k not in ("eof_address", "root_header", "root_scratch")}
    out = {"superblock": sb, "names": f.group_entries}
    for name, d in f.datasets.items():
        lay = {k: v for k, v in d["layout"].items() if k not in ("btree", "addr")}
        out[name] = {"messages": d["messages"], "dataspace": d["dataspace"], "datatype": d["datatype"], "fill": d.get("fill"),
                     "filters": d.get("filters"), "layout": lay,
                     "chunk_offsets": [c["offset"] for c in d.get("chunks", [])],
                     "chunk_filter_masks": [c["filter_mask"] for c in d.get("chunks", [])]}
    return out
