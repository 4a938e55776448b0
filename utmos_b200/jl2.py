"""`.jl` v2: row-compressed presence rows ("both axis pack", SURVEY.md section 8 "next" item 3).

The reference's README (README.md:59-63) floats a second packing step that would take the 1kGP chr22 `.jl` from 29 MB
to 12 MB and leaves it unimplemented "since the overhead it requires slows runtime".  Here the overhead is a GPU kernel:
a row is stored either as its ``np.packbits`` bytes, or -- when that is shorter, i.e. for every rare variant -- as the
sample indices of its carriers; ``utmos_append_packed2`` rebuilds the packed rows in HBM (csrc/ingest.cu
``unpack_rows2_kernel``), so the host never unpacks and PCIe carries the compressed bytes.

On disk a v2 `.jl` is the dict of utmos/convert.py:80-87 with ``GT`` replaced by ``GT2 = {"payload": uint8[...],
"lengths": uint32[V], "idx_bytes": 2 | 4, "n_samples": S}`` (``with_offsets`` adds the running byte offsets the
C ABI takes).  ``utmos convert --pack2`` writes it, ``utmos select``
reads both kinds.  The reference itself cannot read v2 files.
"""
import numpy as np

BLOCK = 8192


def idx_bytes_for(n_samples):
    return 2 if n_samples <= 65536 else 4


def encode(gt_packed, n_samples):
    """uint8 [V, ceil(S/8)] MSB-first rows -> GT2 dict."""
    gt_packed = np.ascontiguousarray(gt_packed, dtype=np.uint8)
    n_rows, pitch = gt_packed.shape
    if pitch != (n_samples + 7) // 8:
        raise ValueError("row pitch must be ceil(S/8)")
    ib = idx_bytes_for(n_samples)
    dt = np.dtype("<u2") if ib == 2 else np.dtype("<u4")
    pieces, lengths = [], np.zeros(n_rows, dtype=np.uint64)
    for r0 in range(0, n_rows, BLOCK):
        block = gt_packed[r0:r0 + BLOCK]
        dense = np.unpackbits(block, axis=1, count=n_samples)
        counts = dense.sum(axis=1, dtype=np.int64)
        sparse = counts * ib < pitch
        rows, cols = np.nonzero(dense)                        # row-major: the indices of a row are contiguous and ascending
        starts = np.concatenate([[0], np.cumsum(counts)])
        for i in range(block.shape[0]):
            if sparse[i]:
                pieces.append(cols[starts[i]:starts[i + 1]].astype(dt).view(np.uint8))
                lengths[r0 + i] = counts[i] * ib
            else:
                pieces.append(block[i])
                lengths[r0 + i] = pitch
    payload = np.concatenate(pieces) if pieces else np.zeros(0, dtype=np.uint8)
    return with_offsets({"payload": payload, "lengths": lengths.astype(np.uint32), "idx_bytes": ib, "n_samples": int(n_samples)})


def with_offsets(gt2):
    """Add ``offsets`` (uint64 [V + 1], exclusive running sum of ``lengths``) if the dict does not carry them yet."""
    if "offsets" not in gt2:
        offsets = np.zeros(len(gt2["lengths"]) + 1, dtype=np.uint64)
        np.cumsum(np.asarray(gt2["lengths"], dtype=np.uint64), out=offsets[1:])
        gt2 = dict(gt2, offsets=offsets)
    return gt2


def for_file(gt2):
    """What goes into the `.jl`: everything but the derived offsets."""
    return {k: v for k, v in gt2.items() if k != "offsets"}


def decode(gt2):
    """GT2 dict -> uint8 [V, ceil(S/8)] (NumPy restatement of unpack_rows2_kernel; tests and the hdf5 writer use it)."""
    gt2 = with_offsets(gt2)
    n_samples, ib = int(gt2["n_samples"]), int(gt2["idx_bytes"])
    pitch = (n_samples + 7) // 8
    offsets = np.asarray(gt2["offsets"], dtype=np.uint64)
    payload = np.asarray(gt2["payload"], dtype=np.uint8)
    n_rows = len(offsets) - 1
    out = np.zeros((n_rows, pitch), dtype=np.uint8)
    dt = np.dtype("<u2") if ib == 2 else np.dtype("<u4")
    for r in range(n_rows):
        b, e = int(offsets[r]), int(offsets[r + 1])
        if e - b == pitch:
            out[r] = payload[b:e]
        else:
            idx = payload[b:e].view(dt).astype(np.int64)
            np.bitwise_or.at(out[r], idx >> 3, (0x80 >> (idx & 7)).astype(np.uint8))
    return out


def n_rows(gt2):
    return len(gt2["lengths"])


def slice_rows(gt2, begin, end):
    """Rows [begin, end) of a GT2 dict (a view: offsets stay absolute into the shared payload)."""
    gt2 = with_offsets(gt2)
    return {"payload": gt2["payload"], "offsets": np.asarray(gt2["offsets"])[begin:end + 1],
            "lengths": np.asarray(gt2["lengths"])[begin:end], "idx_bytes": gt2["idx_bytes"], "n_samples": gt2["n_samples"]}
