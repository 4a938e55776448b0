"""
Select fewest samples with maximum number of variants

Host-side mirror of utmos/select.py: same CLI, same ``load_files`` / ``run_selection`` / ``select_main``
entry points, same report.  What differs is where the numbers are made: ``load_files`` streams the packed
``.jl`` rows / hdf5 chunks into HBM (``_native.DeviceMatrix``) instead of building a dense NumPy matrix, and
``run_selection`` iterates the CUDA greedy loop (csrc/select.cu) instead of ``greedy_select`` /
``calculate_scores`` (utmos/select.py:24-137).  There is no CPU path.
"""
import argparse
import json
import logging
import os
import sys

import joblib
import numpy as np
import pandas as pd

from utmos_b200 import _native, h5lite, jl2
from utmos_b200.convert import read_vcf
from utmos_b200.logutil import setup_logging

MAXMEM = 2  # in GB; accepted for CLI compatibility (utmos/select.py:18) -- the matrix lives in HBM
STEP_BATCH = 4096  # report rows fetched per C-ABI call (rows are still written and flushed one by one)


####################
# Setup/Management #
####################
def resolve_select_count(select_count, num_samples):
    """utmos/select.py:157-159"""
    if select_count < 0:
        return num_samples
    return max(1, int(num_samples * select_count) if select_count < 1 else int(select_count))


def run_selection(data, select_count=0.02, subset=None, exclude=None, weights=None, resume=None):
    """
    Setup the selection calculation and iterate it (utmos/select.py:147-195 + :69-112)
    if select_count [0,1], select that percent of samples
    if select_count >= 1, select that number of samples

    Generator of [sample, var_count, new_count, tot_captured, pct_captured].

    ``resume`` (not in the reference): path of a checkpoint file.  When it exists and belongs to this matrix and these
    options, the rows it holds are yielded again and the selection continues after them; the file is rewritten after
    every batch of picks, so a killed `--count -1` run over a large cohort loses at most one batch.
    """
    matrix = data["data"]
    num_vars, num_samples = matrix.shape
    logging.info("Sample Count %d", num_samples)
    logging.info("Variant Count %d", num_vars)

    select_count = resolve_select_count(select_count, num_samples)
    logging.info("Selecting %d samples", select_count)

    vcf_samples = np.asarray(data["samples"]).astype(str)

    # 1 = can use, 0 = mask, 2 = exclude   (utmos/select.py:168-179)
    sample_mask = np.ones(num_samples, dtype="uint8")
    if subset:
        sample_mask = np.where(np.isin(vcf_samples, subset), 1, 2)
        logging.info("Subsetting to %d samples", len(subset))
    if exclude:
        sample_mask = np.where(np.isin(vcf_samples, exclude), 2, sample_mask)
        logging.info("Excluding %d samples", len(exclude))
    if subset and exclude:
        remain = len(sample_mask) - (sample_mask == 1).sum()
        logging.info("Ending with %d samples", remain)

    sample_weights = None
    if weights is not None:                                   # utmos/select.py:181-187
        logging.info("Setting %d weights", len(weights))
        column = weights["weight"] if hasattr(weights, "columns") else weights
        sample_weights = np.ones(num_samples)
        for pos, i in enumerate(vcf_samples):
            if i in column.index:
                sample_weights[pos] = column.loc[i]

    total_variant_count = np.asarray(data["var_count"][:])
    sample_mask = sample_mask.astype(np.uint8)
    done = None
    if resume:
        if not hasattr(matrix, "export_state"):
            logging.error("--resume is not available for a matrix sharded over several GPUs")
            sys.exit(1)
        tag = _checkpoint_tag(sample_mask, sample_weights, total_variant_count, num_vars)
        state = _load_checkpoint(resume, tag)
        if state is not None:
            logging.info("Resuming after %d picks from %s", len(state["idx"]), resume)
            matrix.import_state(state, sample_weights)
            done = (state["idx"], state["new"], state["stop"])
        else:
            matrix.begin(sample_mask, sample_weights)
        return _greedy_rows(matrix, total_variant_count, select_count, vcf_samples, num_vars, done, (resume, tag))
    matrix.begin(sample_mask, sample_weights)
    return _greedy_rows(matrix, total_variant_count, select_count, vcf_samples, num_vars)


def _checkpoint_tag(sample_mask, sample_weights, var_count, num_vars):
    """Fingerprint of what a checkpoint belongs to: the matrix (shape, var_count) and the scoring options."""
    import hashlib  # pylint: disable=import-outside-toplevel
    h = hashlib.sha256()
    h.update(np.int64(num_vars).tobytes())
    h.update(np.ascontiguousarray(var_count, dtype=np.int64).tobytes())
    h.update(np.ascontiguousarray(sample_mask, dtype=np.uint8).tobytes())
    h.update(b"w" + (np.ascontiguousarray(sample_weights, dtype=np.float64).tobytes() if sample_weights is not None else b""))
    return h.hexdigest()


def _load_checkpoint(path, tag):
    if not os.path.exists(path):
        return None
    with np.load(path, allow_pickle=False) as z:
        if str(z["tag"]) != tag:
            logging.warning("%s belongs to another matrix or other options: starting over", path)
            return None
        return {k: z[k] for k in ("mask", "live", "idx", "new", "score")} | {"tot": int(z["tot"]), "stop": int(z["stop"])}


def _save_checkpoint(path, tag, state):
    tmp = path + ".tmp.npz"
    np.savez(tmp, tag=np.array(tag), **state)
    os.replace(tmp, path)                                    # the old checkpoint stays valid until the new one is complete


def _greedy_rows(matrix, total_variant_count, select_count, vcf_samples, num_vars, done=None, checkpoint=None):
    """Report rows of utmos/select.py:91-112, fed by batches of GPU steps.  ``done`` = (idx, new, stop) of the rows a
    checkpoint already holds: they are yielded first, from the same expressions."""
    tot_captured = 0
    emitted = 0
    first = True
    while emitted < select_count:
        resumed_batch = first and done is not None
        if resumed_batch:
            idx, new, stop = done[0][:select_count], done[1][:select_count], done[2]
        else:
            idx, new, _score, stop = matrix.steps(min(STEP_BATCH, select_count - emitted))
            if checkpoint is not None and len(idx):
                _save_checkpoint(checkpoint[0], checkpoint[1], matrix.export_state())
        first = False
        for use_sample, new_variant_count in zip(idx, new):
            tot_captured += new_variant_count               # np.int64, like counts[use_sample]
            emitted += 1
            yield [
                vcf_samples[use_sample],
                int(total_variant_count[use_sample]),
                int(new_variant_count),
                int(tot_captured),
                round(tot_captured / num_vars, 4)
            ]
        if stop == _native.STOP_ZERO:
            # Backwards compatibility for data without max-alt-af convert
            logging.warning("Ran out of new variants (multi-allelics)")
            return
        if stop == _native.STOP_ALL:
            logging.warning("Ran out of new variants")
            return
        if len(idx) == 0 and not resumed_batch:
            return


class LoadedData(dict):
    """What load_files returns: mapping with 'samples', 'data' (DeviceMatrix) and 'var_count'."""

    def close(self):
        if "data" in self and hasattr(self["data"], "close"):
            self["data"].close()


def _new_matrix(num_samples, af_mode, rows_hint, device, flags, comm):
    """One GPU: DeviceMatrix.  Under torchrun (comm given): this rank's shard of the rows (distributed.ShardedMatrix)."""
    if af_mode == _native.AF_NONE:
        flags &= ~_native.F_REF_TIES                          # count mode has no float sums to replay
    if comm is None:
        return _native.DeviceMatrix(num_samples, af_mode, rows_hint=rows_hint, device=device, flags=flags)
    from utmos_b200.distributed import ShardedMatrix  # pylint: disable=import-outside-toplevel
    return ShardedMatrix(num_samples, af_mode, rows_hint=rows_hint // comm.world + 1, device=device, flags=flags, comm=comm)


def _my_rows(n_rows, comm):
    """[begin, end) of the rows of one input part this rank ingests (all of them on one GPU)."""
    if comm is None:
        return 0, n_rows
    from utmos_b200.distributed import shard_bounds  # pylint: disable=import-outside-toplevel
    return shard_bounds(n_rows, comm.rank, comm.world)


def _load_hdf5(path, device, flags, comm=None):
    """Stream an existing utmos hdf5 (utmos/select.py:250-251) chunk by chunk into HBM."""
    with h5lite.H5File(path) as h5:
        dset = h5["data"]
        num_rows, num_samples = dset.shape
        is_float = dset.dtype != np.dtype(bool)
        if is_float and dset.dtype != np.float32:
            raise h5lite.H5FormatError(f"unexpected data dtype {dset.dtype}")
        af_mode = _native.AF_F32 if is_float else _native.AF_NONE
        matrix = _new_matrix(num_samples, af_mode, num_rows, device, flags, comm)
        table = dset.chunk_table()
        if table is not None and len(table[0]) * dset.chunks[0] >= num_rows:
            # native chunk streamer: pread + LZF decode on all host cores into pinned staging, overlapped with the
            # H2D copies and the packing kernels; under torchrun every rank streams its own run of chunks
            c_begin, c_end = _my_rows(len(table[0]), comm)
            rows_here = max(0, min(num_rows - c_begin * dset.chunks[0], (c_end - c_begin) * dset.chunks[0]))
            matrix.append_h5_chunks(path, table[0][c_begin:c_end], table[1][c_begin:c_end], table[2][c_begin:c_end],
                                    dset.chunks[0], rows_here, is_float, dset.has_lzf)
        else:
            # holes / unordered chunks: a chunk that was never allocated reads as fill-value rows (all zero) in the
            # reference, and those rows count in num_vars = data.shape[0] (utmos/select.py:153)
            r_begin, r_end = _my_rows(num_rows, comm)
            next_row = r_begin

            def _zero_rows(upto):
                nonlocal next_row
                while next_row < upto:
                    n = min(upto - next_row, max(1, (64 << 20) // max(1, num_samples * dset.dtype.itemsize)))
                    matrix.append_dense(np.zeros((n, num_samples), dtype=dset.dtype))
                    next_row += n

            for first, block in dset.iter_chunks():
                lo, hi = max(first, r_begin), min(first + block.shape[0], r_end)
                if lo < hi:
                    _zero_rows(lo)
                    matrix.append_dense(block[lo - first:hi - first])
                    next_row = hi
            _zero_rows(r_end)
        samples = h5["samples"].read()
        stored_var_count = h5["var_count"].read() if "var_count" in h5 else None
    var_count = matrix.finalize()
    if matrix.shape[0] != num_rows:
        raise h5lite.H5FormatError(f"{path}: ingested {matrix.shape[0]} rows, the dataset has {num_rows}")
    if stored_var_count is not None and not np.array_equal(stored_var_count, var_count):
        logging.warning("var_count stored in %s differs from the data; using the stored values", path)
        var_count = np.asarray(stored_var_count)
    return LoadedData(samples=samples, data=matrix, var_count=var_count)


def load_files(in_files, lowmem=None, buffer=32768, calc_af=False, device=0, flags=0, comm=None):
    """
    Load and concatenate multiple files into one HBM-resident matrix (utmos/select.py:241-321).
    if lowmem is a filename, the concatenated informative rows are also written to that hdf5 file
    (bool, or float32 GT*AF when calc_af, utmos/select.py:198-238) so it can be reused later;
    scoring then uses the float32-rounded AF exactly like the reference, which re-reads the file.
    lowmem == 1 means in_files[0] is such an hdf5 file.
    """
    logging.info(f"Loading {len(in_files)} files")
    if lowmem == 1:
        return _load_hdf5(in_files[0], device, flags, comm)

    samples = None
    matrix = None
    writer = None
    load_row_count = 0
    af_mode = _native.AF_NONE
    if calc_af:
        af_mode = _native.AF_F32 if lowmem is not None else _native.AF_F64
    for load_count, i in enumerate(in_files):
        if i.endswith((".vcf.gz", ".vcf")):
            dat = read_vcf(i, lowmem is not None, min(max(1, buffer), 4096), device=device)
        elif i.endswith(".jl"):
            dat = joblib.load(i)
        else:
            logging.error("Unknown filetype %s. Expected `.vcf[.gz]`, `.jl`", i)
            sys.exit(1)

        packed2 = jl2.with_offsets(dat["GT2"]) if "GT" not in dat and "GT2" in dat else None     # `.jl` v2 (jl2.py)
        part_rows = jl2.n_rows(packed2) if packed2 is not None else dat["GT"].shape[0]
        if samples is None:
            samples = np.asarray(dat["samples"]).astype("S")
            matrix = _new_matrix(len(samples), af_mode, part_rows * len(in_files), device, flags, comm)
            if lowmem is not None and (comm is None or comm.rank == 0):
                writer = h5lite.H5Writer(lowmem, samples, float_data=calc_af)     # rank 0 writes the whole file
        r_begin, r_end = _my_rows(part_rows, comm)                                # under torchrun: this rank's rows
        part_af = np.asarray(dat["AF"])[r_begin:r_end] if calc_af else None
        if packed2 is not None:
            matrix.append_packed2(jl2.slice_rows(packed2, r_begin, r_end), part_af)   # decoded on the GPU
        else:
            matrix.append_packed(dat["GT"][r_begin:r_end], part_af)
        if writer is not None:
            writer.append_packed(jl2.decode(packed2) if packed2 is not None else dat["GT"], dat["AF"])
        load_row_count += part_rows
        logging.debug("Loaded %d of %d (%.2f%%) with %d vars", load_count + 1, len(in_files),
                      (load_count + 1) / len(in_files) * 100, load_row_count)

    var_count = matrix.finalize()
    logging.debug("Average of %d variants per-sample", np.mean(var_count) if len(var_count) else 0)
    if writer is not None:
        writer.close(var_count)
    return LoadedData(samples=samples, data=matrix, var_count=var_count)


###################
# Input utilities #
###################
def parse_sample_lists(argument):
    """
    --subset / --exclude values -> flat list of sample names (contract of utmos/select.py:327-340): an item that
    names an existing file contributes one name per line, any other item is split at commas.
    """
    names = []
    for item in argument or ():
        if os.path.exists(item):
            with open(item, "r") as fh:
                names += [line.strip() for line in fh]
        else:
            names += item.split(",")
    return names


def parse_weights(argument):
    """
    --weights TSV (sample <tab> weight, no header) -> DataFrame indexed by sample with one column ``weight``
    (the shape run_selection expects, utmos/select.py:343-352); None when the option was not given.
    """
    if not argument:
        return None
    names, values = [], []
    with open(argument, "r") as fh:
        for lineno, line in enumerate(fh, 1):
            line = line.rstrip("\r\n")
            if not line:
                continue
            fields = line.split("\t")
            if len(fields) != 2:
                raise ValueError(f"{argument}:{lineno}: expected `sample<TAB>weight`")
            names.append(fields[0])
            values.append(float(fields[1]))
    weights = np.asarray(values, dtype=np.float64)
    if len(weights) and np.all(weights == np.floor(weights)):
        weights = weights.astype(np.int64)                    # integers stay integers, like a csv reader infers
    return pd.DataFrame({"weight": weights}, index=pd.Index(names, name="sample"))


# option table of `utmos select` (flags, defaults and meaning of utmos/select.py:359-397, plus --device)
_SELECT_OPTIONS = (
    ("Selection", (
        (("-c", "--count"), dict(type=float, default=0.02,
                                 help="how many samples: a fraction of the cohort below 1, a number from 1 up, -1 for all (%(default)s)")),
        (("-o", "--out"), dict(type=str, default="/dev/stdout", help="report file (%(default)s)")),
        (("--debug",), dict(action="store_true", help="log at debug level")),
        (("--resume",), dict(type=str, default=None,
                             help="checkpoint file (not in the reference): continue from it if it exists, rewrite it after every batch of picks")),
    )),
    ("Scoring", (
        (("--af",), dict(action="store_true", help="score a variant by its allele frequency instead of 1")),
        (("--exact-ties",), dict(action="store_true",
                                 help="with --af: order samples whose exact scores tie by sample index (exact fixed-point sums; the "
                                      "select loop is ~10x faster) instead of replaying the reference's sequential float64 sums, "
                                      "which is the default because it reproduces the reference's order (not in the reference)")),
        (("--ref-ties",), dict(action="store_true", help="accepted for explicitness: the default behaviour of --af on one GPU")),
        (("--weights",), dict(type=str, default=None, help="TSV of sample<TAB>weight; scores are multiplied by it")),
        (("--subset",), dict(type=str, default=None, action="append",
                             help="only these samples can be picked: a file of names or a comma separated list (repeatable)")),
        (("--exclude",), dict(type=str, default=None, action="append",
                              help="these samples are never picked: a file of names or a comma separated list (repeatable)")),
    )),
    ("Memory", (
        (("--lowmem",), dict(type=str, default=None, help="hdf5 file to create from the inputs, or to read when it is the only input")),
        (("--buffer",), dict(type=int, default=32768, help="variants per block while reading VCFs (%(default)s)")),
        (("--maxmem",), dict(type=int, default=2, help="accepted for compatibility: the matrix lives in HBM (%(default)s)")),
    )),
    ("Device", (
        (("--device",), dict(type=int, default=int(os.environ.get("UTMOS_DEVICE", "0")), help="CUDA device index (%(default)s)")),
    )),
)


def _fail(message, *fmt):
    logging.error(message, *fmt)
    sys.exit(1)


def parse_args(args):
    """
    Command line of `utmos select` (options and input rules of utmos/select.py:355-418)
    """
    parser = argparse.ArgumentParser(prog="select", description="Select fewest samples with maximum number of variants")
    parser.add_argument("in_files", nargs="*", type=str, help="VCF (.vcf, .vcf.gz), .jl or one .hdf5 input")
    for title, options in _SELECT_OPTIONS:
        group = parser.add_argument_group(f"{title} options")
        for flags, spec in options:
            group.add_argument(*flags, **spec)
    args = parser.parse_args(args)
    setup_logging(args.debug)

    # input rules: an hdf5 file stands alone; without inputs --lowmem names the hdf5 to read; an hdf5 input is
    # streamed (lowmem = 1 marks "in_files[0] is an existing utmos hdf5")
    n_hdf5 = sum(name.endswith(".hdf5") for name in args.in_files)
    if n_hdf5 and len(args.in_files) > 1:
        _fail("Cannot provide hdf5 with multiple input files")
    if not args.in_files:
        if not args.lowmem:
            _fail("No input files provided")
        args.in_files, args.lowmem = [args.lowmem], 1
    elif n_hdf5 and not args.lowmem:
        logging.info("Switching on lowmem for hdf5 input")
        args.lowmem = 1
    logging.info("Params:\n%s", json.dumps(vars(args), indent=4))
    return args


def _torchrun_comm(args):
    """Under `torchrun -m utmos_b200 select ...`: host collectives of this rank (rows sharded over the ranks, SURVEY.md
    8e); rank 0 keeps the report, the others write theirs to the null device.  None on a single process."""
    if int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return None
    import torch.distributed as dist  # pylint: disable=import-outside-toplevel
    from utmos_b200.distributed import HostCollectives  # pylint: disable=import-outside-toplevel
    if not dist.is_initialized():
        dist.init_process_group("gloo")
    comm = HostCollectives()
    args.device = int(os.environ.get("LOCAL_RANK", str(comm.rank)))
    if comm.rank != 0:
        args.out = os.devnull
    return comm


REPORT_COLUMNS = ("sample", "var_count", "new_count", "tot_captured", "pct_captured")


def select_main(cmdargs):
    """
    `utmos select`: load, check the data against --af, stream the report (utmos/select.py:421-448)
    """
    global MAXMEM  # pylint: disable=global-statement
    args = parse_args(cmdargs)
    comm = _torchrun_comm(args)
    # --af: the reference's own order at exact-arithmetic ties unless --exact-ties; several GPUs
    # only have the exact order
    ref_ties = not args.exact_ties
    if ref_ties and comm is not None:
        if args.ref_ties:
            _fail("--ref-ties is not available on several GPUs")
        if args.af:
            logging.warning("several GPUs: --af ties are ordered by exact sums and sample index (--exact-ties)")
        ref_ties = False
    flags = _native.F_REF_TIES if ref_ties else 0           # only AF matrices look at it (load_files)
    data = load_files(args.in_files, args.lowmem, args.buffer, args.af, device=args.device, flags=flags, comm=comm)
    stored_af = data["data"].dtype != bool                   # float data = GT * AF, made with --af
    if stored_af and flags and hasattr(data["data"], "info") and not data["data"].info().get("ref_ties", 1):
        logging.warning("no sample-major copy of the matrix in HBM: ties of --af scores are ordered by exact sums and "
                        "sample index (as with --exact-ties), not by the reference's float64 accumulation order")
    if args.af and not stored_af:
        logging.critical("HDF5 file doesn't appear to be created with --af weighted scores, remove --af or recreate hdf5")
        sys.exit(1)
    if stored_af and not args.af:                             # the reference only warns here and carries on (:432-433)
        logging.critical("HDF5 file appears to be created with --af weighted scores, add --af or recreate hdf5")
    MAXMEM = args.maxmem
    subset, exclude = parse_sample_lists(args.subset), parse_sample_lists(args.exclude)
    weights = parse_weights(args.weights)
    with open(args.out, "w") as report:
        report.write("\t".join(REPORT_COLUMNS) + "\n")
        for row in run_selection(data, args.count, subset, exclude, weights, resume=args.resume):
            logging.info("Selected %s (%.1f%% of variants)", row[0], row[4] * 100)
            report.write("\t".join(str(value) for value in row) + "\n")
            report.flush()                                    # one line per pick, visible as soon as it is made
    data.close()
    logging.info("Finished utmos")
