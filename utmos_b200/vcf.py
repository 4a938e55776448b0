"""Host-side VCF genotype reader: text -> (samples, int8 GT[V, S, 2]).

Stands in for ``allel.read_vcf(fields=["calldata/GT", "samples"])`` (utmos/convert.py:50-53, scikit-allel
1.3.5): genotypes are parsed to a fixed ploidy-2 int8 tensor, a missing allele ('.', or an absent second
allele of a haploid call) is -1, phasing is ignored.  Text parsing stays on the host (SURVEY.md section 2
row 3); the numeric work on the tensor happens in the K1 kernel (csrc/convert.cu).
"""
import gzip

import numpy as np


def _open(path):
    if path.endswith(".gz"):
        return gzip.open(path, "rt")
    return open(path, "r")


def _parse_gt(token):
    """'0|1' / '1/2' / './.' / '1' -> (a0, a1)."""
    sep = token.find("|")
    if sep < 0:
        sep = token.find("/")
    if sep < 0:
        first, second = token, "."
    else:
        first, second = token[:sep], token[sep + 1:]
        nxt = second.find("|")
        if nxt < 0:
            nxt = second.find("/")
        if nxt >= 0:
            second = second[:nxt]                   # ploidy > 2 is truncated like allel's numbers=2
    a0 = int(first) if first.isdigit() else -1
    a1 = int(second) if second.isdigit() else -1
    return a0, a1


SLAB_BYTES = 256 << 20          # uncompressed text handled at a time (a 1kGP chromosome VCF is ~11 GB of text)


def _bgzf_run(buf, slab_bytes):
    """(end, size): the longest run of whole BGZF blocks buf[:end] that inflates to about slab_bytes (``size`` bytes;
    end 0: no complete block yet), or None when the bytes are not BGZF blocks."""
    pos, size, n = 0, 0, len(buf)
    while pos < n and size < slab_bytes:
        if pos + 18 > n:
            break                                          # header incomplete: wait for more bytes
        if buf[pos:pos + 3] != b"\x1f\x8b\x08" or not buf[pos + 3] & 4:
            return None
        xlen = buf[pos + 10] | (buf[pos + 11] << 8)
        q, xend, bsize = pos + 12, pos + 12 + xlen, -1
        if xend > n:
            break
        while q + 4 <= xend:
            slen = buf[q + 2] | (buf[q + 3] << 8)
            if buf[q:q + 2] == b"BC" and slen == 2:
                bsize = (buf[q + 4] | (buf[q + 5] << 8)) + 1
            q += 4 + slen
        if bsize < 0:
            return None
        if pos + bsize > n:
            break
        size += int.from_bytes(buf[pos + bsize - 4:pos + bsize], "little")
        pos += bsize
    return pos, size


def _text_pieces(fh, threads, slab_bytes):
    """Uncompressed text of an open binary stream (a file or a pipe such as /dev/stdin, README.md:80-84 of the
    reference), about slab_bytes at a time, so memory stays bounded whatever the size of the VCF."""
    import zlib  # pylint: disable=import-outside-toplevel
    from utmos_b200 import _native  # pylint: disable=import-outside-toplevel
    read = max(1 << 16, slab_bytes // 4)
    buf = fh.read(read)
    if buf[:2] != b"\x1f\x8b":                             # plain text
        while buf:
            yield np.frombuffer(buf, dtype=np.uint8)
            buf = fh.read(slab_bytes)
        return
    run = _bgzf_run(buf, 1)
    while run is not None and run[0] == 0:                 # first block not complete yet
        more = fh.read(read)
        if not more:
            break
        buf += more
        run = _bgzf_run(buf, 1)
    if run is not None and run[0]:                         # BGZF: runs of whole blocks, inflated in parallel
        eof = False
        while buf:
            run = _bgzf_run(buf, slab_bytes)
            while run is not None and run[1] < slab_bytes and not eof:
                more = fh.read(read)
                eof = not more
                buf += more
                run = _bgzf_run(buf, slab_bytes)
            if run is None:
                raise ValueError("BGZF stream continues with bytes that are not a BGZF block")
            if run[0] == 0:
                raise ValueError("truncated BGZF block at the end of the input")
            yield _native.gz_inflate(memoryview(buf)[:run[0]], threads)
            buf = buf[run[0]:]
    else:                                                  # plain gzip (possibly several members): one sequential stream
        dec = zlib.decompressobj(wbits=31)
        out, held = [], 0
        while buf:
            data = buf
            while data:
                if dec.eof:                                # next member
                    dec = zlib.decompressobj(wbits=31)
                out.append(dec.decompress(data, max(1, slab_bytes - held)))
                held += len(out[-1])
                data = dec.unused_data if dec.eof else dec.unconsumed_tail
                if held >= slab_bytes:
                    yield np.frombuffer(b"".join(out), dtype=np.uint8)
                    out, held = [], 0
            buf = fh.read(read)
        if not dec.eof:
            raise ValueError("truncated gzip stream")
        if held:
            yield np.frombuffer(b"".join(out), dtype=np.uint8)


def _text_slabs(path, threads, slab_bytes):
    """Yield (uint8 text, is_last) pieces of the uncompressed file, in order (one piece of look-ahead)."""
    with open(path, "rb") as fh:
        prev = None
        for piece in _text_pieces(fh, threads, slab_bytes):
            if prev is not None:
                yield prev, False
            prev = piece
        yield (prev if prev is not None else np.zeros(0, dtype=np.uint8)), True


def read_vcf_genotypes(path, chunk_length=2000, threads=0, slab_bytes=SLAB_BYTES):
    """Yield (samples ndarray[str], int8 GT [n<=chunk_length, S, 2]) blocks in file order.

    The text work is native (csrc/vcfio.cu): BGZF blocks are inflated in parallel, a slab of text at a time, and the
    data lines are tokenised by all host cores straight into the int8 tensor.  ``read_vcf_genotypes_py`` below is
    the pure-Python restatement of the same semantics that the tests pin the native path to."""
    from utmos_b200 import _native  # pylint: disable=import-outside-toplevel
    samples = None
    carry = np.zeros(0, dtype=np.uint8)
    emitted = False
    for slab, is_last in _text_slabs(path, threads, slab_bytes):
        text = np.concatenate([carry, slab]) if len(carry) else slab
        offset = 0
        if samples is None:
            blob = text.tobytes()
            pos = 0 if blob.startswith(b"#CHROM") else blob.find(b"\n#CHROM") + 1
            if pos == 0 and not blob.startswith(b"#CHROM"):
                if is_last:
                    raise ValueError(f"{path}: no #CHROM header line")
                carry = text                               # header longer than one slab: keep reading
                continue
            end = blob.find(b"\n", pos)
            if end < 0:
                if not is_last:
                    carry = text
                    continue
                end = len(blob)
            samples = np.array(blob[pos:end].rstrip(b"\r").decode().split("\t")[9:])     # CRLF files: no '\r' in the last name
            if len(samples) == 0:
                raise ValueError(f"{path}: no sample columns")
            offset = min(end + 1, len(text))
            del blob
        while offset < len(text):
            gts, used = _native.vcf_parse_gt(text, offset, len(samples), chunk_length, is_last, threads)
            offset += used
            if gts.shape[0]:
                emitted = True
                yield samples, gts
            if used == 0:
                break
        carry = text[offset:].copy()                       # an incomplete last line waits for the next slab
    if samples is None:
        raise ValueError(f"{path}: no #CHROM header line")
    if not emitted:
        yield samples, np.zeros((0, len(samples), 2), dtype=np.int8)


def read_vcf_genotypes_py(path, chunk_length=2000):
    """Pure-Python restatement (test infrastructure for the native tokenizer; slow: ~2.6 M genotypes/s)."""
    samples = None
    rows = []
    cache = {}
    with _open(path) as fh:
        for line in fh:
            if line.startswith("##"):
                continue
            if line.startswith("#"):
                samples = np.array(line.rstrip("\r\n").split("\t")[9:])
                continue
            fields = line.rstrip("\r\n").split("\t")
            fmt = fields[8].split(":")
            try:
                gi = fmt.index("GT")
            except ValueError:
                gi = -1
            row = np.full((len(samples), 2), -1, dtype=np.int8)
            if gi >= 0:
                for s, call in enumerate(fields[9:]):
                    if gi == 0 and ":" not in call:
                        tok = call
                    else:
                        parts = call.split(":")
                        tok = parts[gi] if gi < len(parts) else "."     # trailing subfields may be dropped (VCF spec)
                    got = cache.get(tok)
                    if got is None:
                        got = _parse_gt(tok)
                        cache[tok] = got
                    row[s, 0] = got[0]
                    row[s, 1] = got[1]
            rows.append(row)
            if len(rows) >= chunk_length:
                yield samples, np.stack(rows)
                rows = []
    if samples is None:
        raise ValueError(f"{path}: no #CHROM header line")
    if rows:
        yield samples, np.stack(rows)
    elif samples is not None and not rows:
        yield samples, np.zeros((0, len(samples), 2), dtype=np.int8)
