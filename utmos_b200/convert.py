"""convert vcfs to lightweight numpy arrays

Host-side mirror of utmos/convert.py (same CLI, same ``read_vcf`` return value); the numeric core -- presence,
het/hom totals, max-alt allele frequency, bit packing, singleton flags -- runs in the K1 CUDA kernel
(csrc/convert.cu) through ``_native.convert_gt``.
"""
import argparse
import json
import logging

import joblib
import numpy as np

from utmos_b200 import _native, jl2
from utmos_b200.logutil import setup_logging
from utmos_b200.vcf import read_vcf_genotypes


# option table of `utmos convert` (flags and defaults of utmos/convert.py:20-35)
_CONVERT_OPTIONS = (
    (("in_file",), dict(type=str, help="VCF to read (.vcf, .vcf.gz, or /dev/stdin)")),
    (("out_file",), dict(type=str, help=".jl file to write")),
    (("--no-singleton",), dict(action="store_true", help="drop variants whose allele 0 or allele 1 is seen exactly once")),
    (("--lowmem",), dict(action="store_true", help="accepted for compatibility: the VCF is streamed block by block anyway")),
    (("-B", "--buffer"), dict(type=int, default=50000, help="variants per block (%(default)s)")),
    (("-c", "--compress"), dict(type=int, default=5, help="joblib compression level of the .jl file (%(default)s)")),
    (("--pack2",), dict(action="store_true",
                        help="write the row-compressed .jl v2 (the README's \"both axis pack\": rare variants as carrier lists; "
                             "read by this `utmos select`, not by the reference)")),
)


def parse_args(args):
    """
    Command line of `utmos convert`
    """
    parser = argparse.ArgumentParser(prog="convert", description="VCF genotypes -> bit-packed presence matrix (.jl)")
    for flags, spec in _CONVERT_OPTIONS:
        parser.add_argument(*flags, **spec)
    args = parser.parse_args(args)
    setup_logging()
    logging.info("Params:\n%s", json.dumps(vars(args), indent=4))
    return args


def read_vcf(in_file, lowmem=False, chunk_length=2000, no_singleton=False, device=0):
    """
    Read a vcf's genotypes and return numpy arrays (utmos/convert.py:43-88).

    ``lowmem`` is accepted for compatibility: genotype blocks of ``chunk_length`` variants are always
    streamed through the GPU, so host memory never holds more than one block of int8 genotypes.
    """
    del lowmem
    logging.info("Reading VCF")
    packed_parts, af_parts = [], []
    num_hets = num_homs = 0
    samples = None
    removed = 0
    for samples, gts in read_vcf_genotypes(in_file, chunk_length):
        if gts.shape[0] == 0:
            continue
        # utmos/convert.py:58-62 drops singleton rows BEFORE presence / stats / AF are computed: presence and AF are per row,
        # so the kernel only has to leave those rows out of the het / hom totals; their rows are compacted away here
        packed, af, het, hom, single = _native.convert_gt(gts, device, drop_singletons=no_singleton)
        if no_singleton and single.any():
            removed += int(single.sum())
            packed, af = packed[~single], af[~single]
        num_hets += het
        num_homs += hom
        packed_parts.append(packed)
        af_parts.append(af)
    if samples is None:
        raise ValueError(f"{in_file}: empty VCF")
    if no_singleton:
        logging.info("Removing %d singletons", removed)
    logging.info(f"{num_hets} hets")
    logging.info(f"{num_homs} homs")
    pitch = (len(samples) + 7) // 8
    data = {"samples": samples.astype(str)}
    data["GT"] = np.concatenate(packed_parts) if packed_parts else np.zeros((0, pitch), dtype=np.uint8)
    data["AF"] = np.concatenate(af_parts) if af_parts else np.zeros((0, 1), dtype=np.float64)
    data["stats"] = {"num_het": np.int64(num_hets), "num_hom": np.int64(num_homs)}
    return data


def cvt_main(cmdargs):
    """
    `utmos convert`: one VCF -> one .jl (the dict of utmos/convert.py:80-87, written like :98)
    """
    args = parse_args(cmdargs)
    data = read_vcf(args.in_file, args.lowmem, args.buffer, args.no_singleton)
    rows, n_samples = data["GT"].shape[0], len(data["samples"])
    logging.info("Saving %d variants x %d samples (%d het, %d hom-alt calls)", rows, n_samples,
                 int(data["stats"]["num_het"]), int(data["stats"]["num_hom"]))
    if args.pack2:
        data["GT2"] = jl2.for_file(jl2.encode(data.pop("GT"), n_samples))
        logging.info("Row-compressed rows: %d bytes for %d packed", len(data["GT2"]["payload"]), rows * ((n_samples + 7) // 8))
    joblib.dump(data, args.out_file, compress=args.compress)
    logging.info("Finished conversion")
