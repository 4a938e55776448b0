"""CPU-side tests: the C-ABI library loads and exports its surface, host parsers, hdf5 reader, CLI plumbing."""
import os
import re

import numpy as np
import pytest

from oracle import select_oracle as orc
from tests import helpers as H
from utmos_b200 import _native, h5lite, synth, vcf
from utmos_b200 import select as usel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "utmos_b200.h")).read()
    declared = set(re.findall(r"\b(utmos_[a-z0-9_]+)\s*\(", header))
    declared.discard("utmos_ctx")
    lib = _native.lib()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_native.SYMBOLS)


def test_no_cpu_fallback_without_gpu():
    if _native.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(_native.NativeError) as err:
        _native.DeviceMatrix(8)
    assert err.value.code == _native.E_NOGPU
    with pytest.raises(_native.NativeError):
        _native.convert_gt(np.zeros((2, 3, 2), dtype=np.int8))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "utmos_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, name)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, name


def test_lzf_roundtrip_and_fixture_decode():
    rng = np.random.default_rng(1)
    for n in (0, 1, 2, 3, 17, 300, 5000, 70000):
        for kind in range(3):
            if kind == 0:
                data = rng.integers(0, 2, n, dtype=np.uint8)
            elif kind == 1:
                data = np.repeat(rng.integers(0, 255, max(1, n // 40), dtype=np.uint8), 40)[:n]
            else:
                data = rng.integers(0, 256, n, dtype=np.uint8)
            comp = _native.lzf_compress(data.tobytes())
            if comp is not None:
                assert _native.lzf_decompress(comp, len(data)).tobytes() == data.tobytes()


def test_h5lite_reads_reference_fixtures():
    parts = H.load_jl_parts(["chunk2.jl"])
    dense, _, _ = orc.unpack_and_filter(parts[0]["GT"], 2504)
    with h5lite.H5File(H.fixture("tiny.hdf5")) as h5:
        assert set(h5.keys()) == {"data", "samples", "var_count"}
        data = h5["data"]
        assert data.shape == dense.shape and data.dtype == np.dtype(bool) and data.chunks == (99, 2504)
        assert np.array_equal(data.read(), dense)
        assert np.array_equal(h5["var_count"].read(), dense.sum(axis=0))
        assert list(h5["samples"].read().astype(str)) == list(np.asarray(parts[0]["samples"]).astype(str))
        first_rows = [first for first, _ in data.iter_chunks()]
        assert first_rows == list(range(0, 995, 99))
    parts = H.load_jl_parts(["chunk0.jl", "chunk1.jl"])
    matrix, var_count = orc.load_parts(parts, 2504, float32_af=True)
    with h5lite.H5File(H.fixture("tiny.af.hdf5")) as h5:
        assert h5["data"].dtype == np.float32
        assert np.array_equal(h5["data"].read(), matrix)
        assert np.array_equal(h5["var_count"].read(), var_count)


@pytest.mark.parametrize("name", ["tiny.hdf5", "tiny.af.hdf5"])
def test_h5lite_chunk_table_feeds_the_native_streamer(name):
    """The byte ranges handed to utmos_append_h5_chunks decode to exactly the blocks iter_chunks yields."""
    with h5lite.H5File(H.fixture(name)) as h5:
        dset = h5["data"]
        addr, nbytes, fmask = dset.chunk_table()
        rows, n_samples = dset.chunks
        assert dset.has_lzf and len(addr) * rows >= dset.shape[0] and n_samples == dset.shape[1]
        chunk_bytes = rows * n_samples * dset.dtype.itemsize
        blocks = list(dset.iter_chunks())
        assert len(blocks) == len(addr)
        with open(H.fixture(name), "rb") as fh:
            for (first, block), a, n, m in zip(blocks, addr, nbytes, fmask):
                fh.seek(int(a))
                raw = fh.read(int(n))
                full = np.frombuffer(raw, np.uint8) if m & 1 else _native.lzf_decompress(raw, chunk_bytes)
                got = full.view(dset.dtype).reshape(rows, n_samples)[:block.shape[0]]
                assert np.array_equal(got, block), first


@pytest.mark.parametrize("name", ["chunk0", "chunk1"])
def test_vcf_reader_and_convert_oracle_reproduce_fixture_jl(name):
    """chunkN.vcf.gz -> GT bits and het/hom stats of chunkN.jl (AF: fixture holds the older definition)."""
    gold = H.load_jl_parts([name + ".jl"])[0]
    blocks = list(vcf.read_vcf_genotypes(H.fixture(name + ".vcf.gz"), 400))
    samples = blocks[0][0]
    gts = np.concatenate([b[1] for b in blocks])
    assert gts.shape == (1000, 2504, 2)
    assert list(samples) == list(np.asarray(gold["samples"]).astype(str))
    out = orc.convert_gt_c(gts)
    assert np.array_equal(out["GT"], gold["GT"])
    assert out["stats"]["num_het"] == gold["stats"]["num_het"]
    assert out["stats"]["num_hom"] == gold["stats"]["num_hom"]
    # biallelic rows: max-alt AF == allele-1 AF, which is what the fixture stores
    multi = (gts > 1).any(axis=(1, 2))
    assert np.array_equal(out["AF"][~multi], gold["AF"][~multi])


def test_vcf_gt_token_forms():
    assert vcf._parse_gt("0|1") == (0, 1)
    assert vcf._parse_gt("1/2") == (1, 2)
    assert vcf._parse_gt("./.") == (-1, -1)
    assert vcf._parse_gt(".|1") == (-1, 1)
    assert vcf._parse_gt("1") == (1, -1)
    assert vcf._parse_gt("10|12") == (10, 12)
    assert vcf._parse_gt("0/1/1") == (0, 1)


def _bgzf(data, block=700):
    """Minimal BGZF writer (test helper): independent gzip members with the 'BC' extra subfield + EOF block."""
    import struct
    import zlib
    out = []
    for i in list(range(0, len(data), block)) + [None]:
        chunk = b"" if i is None else data[i:i + block]
        comp = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = comp.compress(chunk) + comp.flush()
        bsize = 12 + 6 + len(body) + 8
        out.append(b"\x1f\x8b\x08\x04" + b"\x00" * 4 + b"\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1)
                   + body + struct.pack("<II", zlib.crc32(chunk), len(chunk)))
    return b"".join(out)


def test_native_gz_inflate_bgzf_plain_and_multimember():
    import gzip
    rng = np.random.default_rng(3)
    data = bytes(rng.integers(0, 4, 50_000, dtype=np.uint8) + 48) + b"tail\n"
    for blob in (_bgzf(data), gzip.compress(data), gzip.compress(data[:1000]) + gzip.compress(data[1000:])):
        assert _native.gz_inflate(blob).tobytes() == data
    assert _native.gz_inflate(_bgzf(b"")).tobytes() == b""
    with pytest.raises(ValueError):
        _native.gz_inflate(b"plain text, not gzip at all")
    broken = bytearray(_bgzf(data))
    broken[40] ^= 0xff
    with pytest.raises(_native.NativeError):
        _native.gz_inflate(bytes(broken))


VCF_HEAD = "##fileformat=VCFv4.2\n##contig=<ID=1>\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tA\tB\tC\tD\n"
VCF_LINES = [
    "1\t10\t.\tA\tC\t.\t.\t.\tGT\t0|1\t1/1\t./.\t0|0",
    "1\t11\t.\tA\tC,G\t.\t.\t.\tGT:DP\t0/2:7\t.|1:3\t1:9\t0/1/2:4",          # haploid call, ploidy 3, '.|1'
    "1\t12\t.\tA\tC\t.\t.\t.\tDP:GT\t7:0|1\t3:1|1\t9\t4:.",                    # GT second; a call without that subfield
    "1\t13\t.\tA\tC\t.\t.\t.\tDP\t7\t3\t9\t4",                                 # no GT in FORMAT -> all missing
    "1\t14\t.\tA\t<X>\t.\t.\t.\tGT\t10|12\t0/x\t|\t/1",                        # two-digit alleles, junk tokens
    "1\t15\t.\tA\tC\t.\t.\t.\tGT\t0|1\t1|0",                                    # fewer sample columns than the header
    "1\t16\t.\tA\tC\t.\t.\t.\tGT\t0/1|1\t1|0/1\t0|1|2\t0",                    # mixed separators
]


@pytest.mark.parametrize("ending", ["\n", "", "\r\n"], ids=["newline", "no_final_newline", "crlf"])
@pytest.mark.parametrize("container", ["plain", "gzip", "bgzf"])
def test_native_vcf_tokenizer_matches_python_restatement(tmp_path, ending, container):
    """csrc/vcfio.cu against vcf.read_vcf_genotypes_py on genotype forms the fixtures do not contain."""
    import gzip
    sep = "\r\n" if ending == "\r\n" else "\n"
    text = (VCF_HEAD.replace("\n", sep) + sep.join(VCF_LINES) + ending).encode()     # CRLF files: header lines too
    path = tmp_path / ("t.vcf" if container == "plain" else "t.vcf.gz")
    path.write_bytes(text if container == "plain" else gzip.compress(text) if container == "gzip" else _bgzf(text, 97))
    want = list(vcf.read_vcf_genotypes_py(str(path), 3))
    for slab in (1 << 20, 150):                                  # 150-byte slabs: lines straddle slab boundaries
        got = list(vcf.read_vcf_genotypes(str(path), 3, threads=3, slab_bytes=slab))
        assert list(got[0][0]) == list(want[0][0]) == ["A", "B", "C", "D"]       # no '\r' glued to the last name
        a, b = np.concatenate([g for _, g in got]), np.concatenate([g for _, g in want])
        assert a.shape == b.shape == (len(VCF_LINES), 4, 2) and np.array_equal(a, b)
        assert all(g.shape[0] <= 3 for _, g in got)


def test_native_vcf_tokenizer_rejects_what_it_cannot_represent(tmp_path):
    bad = VCF_HEAD + "1\t10\t.\tA\tC\t.\t.\t.\tGT\t0|1\t1/1\t./.\t0|0\t1|1\n"          # five calls, four samples
    (tmp_path / "a.vcf").write_text(bad)
    with pytest.raises(_native.NativeError):
        list(vcf.read_vcf_genotypes(str(tmp_path / "a.vcf")))
    big = VCF_HEAD + "1\t10\t.\tA\tC\t.\t.\t.\tGT\t0|200\t1/1\t./.\t0|0\n"                  # allele 200 does not fit int8
    (tmp_path / "b.vcf").write_text(big)
    with pytest.raises(_native.NativeError):
        list(vcf.read_vcf_genotypes(str(tmp_path / "b.vcf")))
    (tmp_path / "c.vcf").write_text("##only meta\n")
    with pytest.raises(ValueError):
        list(vcf.read_vcf_genotypes(str(tmp_path / "c.vcf")))
    (tmp_path / "d.vcf").write_text(VCF_HEAD)                    # header only: one empty block
    blocks = list(vcf.read_vcf_genotypes(str(tmp_path / "d.vcf")))
    assert len(blocks) == 1 and blocks[0][1].shape == (0, 4, 2)


@pytest.mark.parametrize("name", ["chunk0", "chunk1"])
def test_native_vcf_reader_equals_python_reader_on_fixtures(name):
    path = H.fixture(name + ".vcf.gz")
    want = np.concatenate([g for _, g in vcf.read_vcf_genotypes_py(path, 500)])
    for slab in (256 << 20, 90_000):
        got = np.concatenate([g for _, g in vcf.read_vcf_genotypes(path, 333, slab_bytes=slab)])
        assert np.array_equal(got, want)


def test_convert_oracle_edge_genotypes():
    gt = np.array([[[0, 0], [0, 1], [1, 1], [-1, 1], [2, 1], [-1, -1], [2, 2], [1, -1]]], dtype=np.int8)
    out = orc.convert_gt_c(gt)
    bits = np.unpackbits(out["GT"], axis=1, count=8)[0]
    assert list(bits) == [0, 1, 1, 0, 1, 0, 1, 0]
    assert out["stats"] == {"num_het": 2, "num_hom": 2}
    assert out["AF"][0, 0] == 6 / 12                      # allele 1: 6 of 12 called alleles; allele 2: 3
    empty = orc.convert_gt_c(np.full((1, 4, 2), -1, dtype=np.int8))
    assert np.isnan(empty["AF"][0, 0])


def test_convert_oracle_and_vcf_parser_on_hand_built_rows(tmp_path):
    """Missing, half-missing, haploid, AN = 0 and multi-allelic rows (none of them in the reference's fixtures): the host VCF
    parsers and the convert oracle against answers worked out from scikit-allel's definitions (tests/helpers.py)."""
    path = tmp_path / "hand.vcf"
    path.write_text(H.HAND_VCF_HEAD + "\n".join(r[0] for r in H.HAND_VCF_ROWS) + "\n")
    for reader in (vcf.read_vcf_genotypes_py, vcf.read_vcf_genotypes):
        blocks = list(reader(str(path), 100))
        assert list(blocks[0][0]) == ["A", "B", "C", "D"]
        gts = np.concatenate([b[1] for b in blocks])
        out = orc.convert_gt_c(gts)
        bits = np.unpackbits(out["GT"], axis=1, count=4)
        for i, (_line, presence, af, _het, _hom, single) in enumerate(H.HAND_VCF_ROWS):
            assert list(bits[i]) == presence, i
            assert (np.isnan(out["AF"][i, 0]) and np.isnan(af)) or out["AF"][i, 0] == af, i
            assert bool(out["singleton"][i]) == single, i
        assert out["stats"] == {"num_het": sum(r[3] for r in H.HAND_VCF_ROWS), "num_hom": sum(r[4] for r in H.HAND_VCF_ROWS)}


def test_jl2_row_compression_roundtrip():
    """`.jl` v2 (utmos_b200/jl2.py): encode / decode / slice are lossless on the reference's fixtures and on
    ragged sample counts, rare rows become carrier lists and the payload is several times smaller than np.packbits rows."""
    from utmos_b200 import jl2
    part = H.load_jl_parts(["chunk1.jl"])[0]
    g2 = jl2.encode(part["GT"], 2504)
    assert np.array_equal(jl2.decode(jl2.for_file(g2)), part["GT"])
    assert g2["payload"].nbytes * 4 < part["GT"].nbytes and g2["idx_bytes"] == 2
    assert int(g2["lengths"].max()) == 313                                   # common variants stay np.packbits rows
    sl = jl2.slice_rows(g2, 17, 400)
    assert jl2.n_rows(sl) == 383 and int(sl["offsets"][0]) == int(g2["offsets"][17])
    for n_samples in (1, 7, 8, 9, 70001):
        rng = np.random.default_rng(n_samples)
        dense = rng.random((40, n_samples)) < 0.02
        dense[5] = True                                                       # a full row
        dense[6] = False                                                      # an empty row: zero-length entry
        gt = np.packbits(dense, axis=1)
        g = jl2.encode(gt, n_samples)
        assert g["idx_bytes"] == (4 if n_samples > 65536 else 2)
        assert np.array_equal(jl2.decode(g), gt) and int(g["lengths"][6]) == 0


def test_count_resolution_and_sample_lists(tmp_path):
    assert usel.resolve_select_count(-1, 2504) == 2504
    assert usel.resolve_select_count(0.02, 2504) == 50
    assert usel.resolve_select_count(0.01, 2504) == 25
    assert usel.resolve_select_count(0, 2504) == 1
    assert usel.resolve_select_count(1.0, 2504) == 1
    assert usel.resolve_select_count(10, 2504) == 10
    for c in (-1, 0, 0.005, 0.02, 0.9999, 1, 1.5, 20, 5000):
        assert usel.resolve_select_count(c, 2504) == orc.resolve_count(c, 2504)
    lst = tmp_path / "names.txt"
    lst.write_text("A\nB \nC\n")
    assert usel.parse_sample_lists([str(lst), "X,Y"]) == ["A", "B", "C", "X", "Y"]
    assert usel.parse_sample_lists(None) == []
    wts = usel.parse_weights(H.fixture("weights.txt"))
    assert float(wts.loc["HG00280", "weight"]) == 4 and float(wts.loc["NA20320", "weight"]) == 10
    assert usel.parse_weights(None) is None


def test_cli_argument_errors(tmp_path):
    for argv in (["a.hdf5", "b.hdf5"], [], ["nothere.txt"]):
        with pytest.raises(SystemExit) as err:
            args = usel.parse_args(argv)
            usel.load_files(args.in_files, args.lowmem, args.buffer, args.af)
        assert err.value.code == 1
    args = usel.parse_args(["x.hdf5"])
    assert args.lowmem == 1
    args = usel.parse_args(["--lowmem", "y.hdf5"])
    assert args.lowmem == 1 and args.in_files == ["y.hdf5"]
    args = usel.parse_args(["--subset", "a", "--subset", "b", "--exclude", "c", "f.jl"])
    assert args.subset == ["a", "b"] and args.exclude == ["c"] and args.count == 0.02


def test_synth_mirror_is_deterministic_and_informative():
    gt1, af1 = synth.mirror_rows(3, 100, 500, 333)
    gt2, af2 = synth.mirror_rows(3, 0, 600, 333)
    assert np.array_equal(gt1, gt2[100:600]) and np.array_equal(af1, af2[100:600])
    bits = np.unpackbits(gt1, axis=1, count=333)
    assert bits.any(axis=1).all()
    assert not np.unpackbits(gt1, axis=1)[:, 333:].any()
    assert ((af1 > 0) & (af1 <= 1)).all()


def test_h5writer_roundtrip_bool_float_and_multilevel_btree(tmp_path):
    parts = H.load_jl_parts(["chunk0.jl", "chunk1.jl"])
    samples = np.asarray(parts[0]["samples"]).astype("S")
    for float_data in (False, True):
        path = str(tmp_path / f"w{int(float_data)}.hdf5")
        writer = h5lite.H5Writer(path, samples, float_data=float_data)
        for part in parts:
            writer.append_packed(part["GT"], part["AF"])
        matrix, var_count = orc.load_parts(parts, 2504, float32_af=float_data)
        writer.close(var_count)
        with h5lite.H5File(path) as h5:
            assert h5["data"].chunks == (99, 2504)                      # utmos/select.py:205
            assert h5["data"].dtype == (np.float32 if float_data else np.dtype(bool))
            assert np.array_equal(h5["data"].read(), matrix)
            assert np.array_equal(h5["var_count"].read(), var_count)
            assert list(h5["samples"].read()) == list(samples)
    # > 64 chunks forces a two-level chunk B-tree; ragged last chunk
    n_samples = 60000
    path = str(tmp_path / "big.hdf5")
    writer = h5lite.H5Writer(path, synth.sample_names(n_samples).astype("S"))
    rng = np.random.default_rng(0)
    blocks = [rng.random((n, n_samples)) < 0.01 for n in (50, 1, 133, 260, 7)]
    for block in blocks:
        writer.append_dense(block)
    full = np.concatenate(blocks)
    writer.close(full.sum(axis=0))
    with h5lite.H5File(path) as h5:
        assert h5["data"].chunks == (4, n_samples) and h5["data"].shape == full.shape
        assert len(h5["data"]._chunk_index()) == (full.shape[0] + 3) // 4 > 64
        assert np.array_equal(h5["data"].read(), full)


@pytest.mark.parametrize("name", ["tiny.hdf5", "tiny.af.hdf5"])
@pytest.mark.parametrize("path_kind", ["dense", "packed"])
def test_h5writer_structure_equals_h5py_fixture(tmp_path, name, path_kind):
    """utmos/select.py:198-238 / utmos_ssshtests.sh:197-235: a file written by H5Writer for the content of the reference's
    h5py-written fixture has the same structure message by message -- superblock fields, group entries, and for every
    dataset the object-header messages in order with their flags, dataspace (dims, max dims), datatype (class, bit
    fields, size, properties), fill value, filter pipeline (LZF id, name, flags, client values), layout (chunk dims) and
    the chunk offsets / filter masks -- and every B-tree / heap / header invariant of the format holds (tests/h5struct.py,
    an independent strict reader).  Only addresses and compressed chunk sizes may differ.  h5py itself is not
    installed here; this is the closest statement available that stock h5py / the reference can re-read the file."""
    from tests import h5struct
    ref = h5struct.File(H.fixture(name))
    with h5lite.H5File(H.fixture(name)) as h5:
        data, samples, vc = h5["data"].read(), h5["samples"].read(), h5["var_count"].read()
    out = str(tmp_path / "w.hdf5")
    is_float = data.dtype != bool
    w = h5lite.H5Writer(out, samples, float_data=is_float)
    if path_kind == "dense":
        w.append_dense(data[:100])
        w.append_dense(data[100:])
    else:                                                  # from packed .jl rows + AF, the way load_files feeds it
        af = data.max(axis=1).astype(np.float64) if is_float else None
        packed = np.packbits(data != 0, axis=1)
        w.append_packed(packed[:150], None if af is None else af[:150])
        w.append_packed(packed[150:], None if af is None else af[150:])
    w.close(vc)
    ours = h5struct.File(out)
    assert h5struct.comparable(ours) == h5struct.comparable(ref)
    with h5lite.H5File(out) as h5:                          # and the content reads back bit for bit
        assert np.array_equal(h5["data"].read(), data) and h5["data"].dtype == data.dtype
        assert np.array_equal(h5["samples"].read(), samples) and np.array_equal(h5["var_count"].read(), vc)


def test_h5struct_rejects_broken_files(tmp_path):
    """The strict reader really checks: a wrong end-of-file address, a broken sibling pointer and an unsorted key fail."""
    from tests import h5struct
    blob = bytearray(open(H.fixture("tiny.hdf5"), "rb").read())
    bad = bytearray(blob)
    bad[40:48] = (len(blob) + 8).to_bytes(8, "little")             # end-of-file address
    (tmp_path / "a.hdf5").write_bytes(bad)
    with pytest.raises(h5struct.Bad):
        h5struct.File(str(tmp_path / "a.hdf5"))
    ref = h5struct.File(H.fixture("tiny.hdf5"))
    root = ref.datasets["data"]["layout"]["btree"]
    bad = bytearray(blob)
    bad[root + 8:root + 16] = (1234).to_bytes(8, "little")         # left sibling of the root must be undefined
    (tmp_path / "b.hdf5").write_bytes(bad)
    with pytest.raises(h5struct.Bad):
        h5struct.File(str(tmp_path / "b.hdf5"))


def test_h5writer_native_chunk_encoder_matches_numpy_path(tmp_path):
    """Native encoder (unpack + LZF on host threads, csrc/hostio.cu) against the NumPy/dense path of the writer:
    identical files, for bool and float32 data, ragged parts, uninformative rows, threads 1 and many, and parts
    that switch between packed and dense appends while a chunk is half full."""
    rng = np.random.default_rng(5)
    n_samples = 1237                                                   # 202 rows per chunk, 3 pad bits per row
    samples = synth.sample_names(n_samples).astype("S")
    parts = []
    for n in (150, 1, 700, 0, 409, 33):
        dense = rng.random((n, n_samples)) < 0.03
        dense[::17] = False                                            # uninformative rows are dropped
        parts.append((np.packbits(dense, axis=1), rng.random(n), dense))
    full = np.concatenate([p[2] for p in parts])
    full = full[full.any(axis=1)]
    for float_data in (False, True):
        files = []
        for tag, threads in (("py", None), ("n1", 1), ("n8", 8)):
            path = str(tmp_path / f"{tag}{int(float_data)}.hdf5")
            writer = h5lite.H5Writer(path, samples, float_data=float_data)
            for gt, af, _ in parts:
                if threads is None:
                    writer.append_packed_py(gt, af)
                else:
                    writer.append_packed(gt, af, threads=threads)
            writer.close(full.sum(axis=0))
            files.append(open(path, "rb").read())
        assert files[0] == files[1] == files[2]
        with h5lite.H5File(str(tmp_path / f"n8{int(float_data)}.hdf5")) as h5:
            assert h5["data"].shape == full.shape
            if float_data:
                af_all = np.concatenate([p[1] for p in parts])[np.concatenate([p[2] for p in parts]).any(axis=1)]
                assert np.array_equal(h5["data"].read(), (full * af_all.reshape(-1, 1)).astype(np.float32))
            else:
                assert np.array_equal(h5["data"].read(), full)
    # packed, dense, packed: the dense append picks up the packed rows still waiting for their chunk
    path = str(tmp_path / "mixed.hdf5")
    writer = h5lite.H5Writer(path, samples)
    writer.append_packed(parts[0][0], None)
    extra = rng.random((300, n_samples)) < 0.02
    extra[:, 0] = True
    writer.append_dense(extra)
    writer.append_packed(parts[2][0], None)
    want = np.concatenate([parts[0][2][parts[0][2].any(axis=1)], extra, parts[2][2][parts[2][2].any(axis=1)]])
    writer.close(want.sum(axis=0))
    with h5lite.H5File(path) as h5:
        assert np.array_equal(h5["data"].read(), want)


def test_vcf_text_is_streamed_from_pipes_and_large_inputs(tmp_path):
    """`bcftools view ... | utmos convert /dev/stdin out.jl` (reference README.md:80-84): the text arrives through a
    pipe, in every container, in bounded pieces; multi-member gzip and BGZF runs that straddle read boundaries."""
    import gzip
    import subprocess
    import sys
    rng = np.random.default_rng(11)
    n_samples, n_lines = 40, 6000                                       # ~1 MB of text: several 64 KB reads
    names = "\t".join(f"s{i}" for i in range(n_samples))
    alleles = rng.integers(0, 3, (n_lines, n_samples, 2))
    lines = ["1\t%d\t.\tA\tC,G\t.\t.\t.\tGT\t%s" % (i + 1, "\t".join(f"{a}|{b}" for a, b in row))
             for i, row in enumerate(alleles)]
    text = ("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + names + "\n" +
            "\n".join(lines) + "\n").encode()
    third = len(text) // 3
    blobs = {"plain": text, "gzip": gzip.compress(text), "bgzf": _bgzf(text, 4000),
             "gzip_members": gzip.compress(text[:third]) + gzip.compress(text[third:2 * third]) + gzip.compress(text[2 * third:])}
    for name, blob in blobs.items():
        path = tmp_path / f"{name}.bin"
        path.write_bytes(blob)
        for slab in (1 << 30, 100_000):
            got = np.concatenate([g for _, g in vcf.read_vcf_genotypes(str(path), 777, slab_bytes=slab)])
            assert np.array_equal(got, alleles), (name, slab)
        pieces = list(vcf._text_slabs(str(path), 2, 100_000))
        assert len(pieces) >= 8 and [last for _, last in pieces] == [False] * (len(pieces) - 1) + [True]
        assert b"".join(p.tobytes() for p, _ in pieces) == text
    # through a real pipe
    code = ("import sys, numpy as np; sys.path.insert(0, %r); from utmos_b200 import vcf; "
            "g = np.concatenate([g for _, g in vcf.read_vcf_genotypes('/dev/stdin', 500, slab_bytes=200_000)]); "
            "sys.stdout.write('%%d %%d' %% (g.shape[0], int(g.astype(np.int64).sum())))" % ROOT)
    for name in ("plain", "bgzf", "gzip"):
        feeder = subprocess.Popen(["cat", str(tmp_path / f"{name}.bin")], stdout=subprocess.PIPE)
        out = subprocess.run([sys.executable, "-c", code], stdin=feeder.stdout, capture_output=True, text=True, timeout=120, check=True)
        feeder.wait()
        assert out.stdout.split() == [str(n_lines), str(int(alleles.sum()))], (name, out.stderr)
    # damaged containers fail loudly
    (tmp_path / "cut.gz").write_bytes(blobs["gzip"][:-20])
    with pytest.raises(ValueError):
        list(vcf.read_vcf_genotypes(str(tmp_path / "cut.gz")))
    (tmp_path / "cut.bgzf").write_bytes(blobs["bgzf"][:-40])
    with pytest.raises(ValueError):
        list(vcf.read_vcf_genotypes(str(tmp_path / "cut.bgzf")))


class OracleMatrix:
    """Stand-in for _native.DeviceMatrix in CPU tests of the host code: the C oracle does the arithmetic."""

    def __init__(self, n_samples, af_mode=0, rows_hint=0, device=0, flags=0):
        self.n_samples, self.af_mode = n_samples, af_mode
        self.parts, self.afs = [], []

    def append_packed(self, gt, af=None):
        self.parts.append(np.asarray(gt))
        self.afs.append(None if af is None else np.asarray(af, dtype=np.float64).reshape(-1))

    def finalize(self):
        packed = np.concatenate(self.parts)
        keep, var_count = orc.filter_rows_c(packed, self.n_samples)
        self.packed = packed[keep]
        self.af = None
        if self.af_mode != _native.AF_NONE:
            self.af = np.concatenate(self.afs)[keep]
            if self.af_mode == _native.AF_F32:
                self.af = self.af.astype(np.float32).astype(np.float64)
        self.num_vars = int(keep.sum())
        return var_count

    @property
    def shape(self):
        return (self.num_vars, self.n_samples)

    @property
    def dtype(self):
        return np.dtype(bool) if self.af_mode == _native.AF_NONE else np.dtype(np.float64)

    def begin(self, mask, weights=None):
        self.result = orc.greedy_c(self.packed, self.n_samples, mask, weights, self.af, self.n_samples)
        self.pos = 0

    def steps(self, max_steps):
        idx, new, score, stop = self.result
        a, b = self.pos, min(len(idx), self.pos + max_steps)
        self.pos = b
        return idx[a:b], new[a:b], score[a:b], (stop if b == len(idx) else 0)

    def close(self):
        pass


JL_KEYED = [
    ("select_intcnt.txt", ["--count", "10", "chunk1.jl"]),
    ("select_floatcnt.txt", ["--count", "0.01", "chunk2.jl"]),
    ("select_first.txt", ["chunk2.jl"]),
    ("select_multi.txt", ["chunk0.jl", "chunk2.jl"]),
    ("select_exclude.txt", ["-c", "20", "--exclude", "NA21117", "chunk0.jl", "chunk1.jl"]),
    ("select_weights.txt", ["-c", "20", "--weights", "weights.txt", "chunk0.jl"]),
    ("select_af.txt", ["-c", "20", "--af", "chunk0.jl", "chunk1.jl"]),
    ("select_weightsaf.txt", ["-c", "5", "--af", "--weights", "weights.txt", "chunk0.jl", "chunk1.jl"]),
    ("select_weights_subset.txt", ["--subset", "subset.txt", "-c", "5", "--weights", "weights.txt", "chunk0.jl"]),
    ("select_af_subset.txt", ["--subset", "subset.txt", "-c", "5", "--af", "chunk0.jl"]),
]


@pytest.mark.parametrize("key,argv", JL_KEYED, ids=[k for k, _ in JL_KEYED])
def test_select_main_host_code_reproduces_answer_keys(key, argv, tmp_path, monkeypatch):
    """The host side of `utmos select` (argument rules, sample lists, weights, report writer -- utmos/select.py:327-448)
    with the C oracle standing in for the device: byte-identical to the reference's answer keys."""
    monkeypatch.setattr(usel._native, "DeviceMatrix", OracleMatrix)
    out = tmp_path / "report.txt"
    usel.select_main([H.fixture(a) if os.path.exists(H.fixture(a)) else a for a in argv] + ["-o", str(out)])
    assert out.read_text() == H.answer_key(key)


def test_cli_dispatcher(capsys):
    from utmos_b200 import __main__ as cli
    assert cli.run([]) == 0
    assert "convert" in capsys.readouterr().err                      # overview on stderr, status 0 (utmos/__main__.py:42-44)
    assert cli.run(["version"]) == 0
    assert capsys.readouterr().out.strip() == "Utmos v2.2.0"
    assert cli.run(["bogus"]) == 2
    with pytest.raises(SystemExit) as err:
        cli.run(["select"])                                          # no inputs: the command's own error, exit 1
    assert err.value.code == 1
