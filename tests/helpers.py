"""Shared helpers for the test-suite (fixture loading, option plumbing)."""
import json
import os

import joblib
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
REF_FILES = os.path.join(GOLD, "ref", "test_files")
REF_KEYS = os.path.join(GOLD, "ref", "answer_key")


# Positions (in tests/golden/random_cases.json) of the recorded `--af` cases whose report differs from the unmodified
# reference's: exactly these 13 of the 72 `--af` cases, each proven a near-tie of the reference's own float64 sums
# (tests/test_oracle.py).  Pinned as a set: a 14th divergence, or one that disappears, fails the suite.
NEAR_TIE_CASES = [16, 26, 71, 86, 128, 131, 143, 146, 166, 168, 171, 173, 178]


def fixture(name):
    return os.path.join(REF_FILES, name)


def answer_key(name):
    with open(os.path.join(REF_KEYS, name)) as fh:
        return fh.read()


def load_jl_parts(names):
    """joblib payloads of reference fixtures, in order."""
    return [joblib.load(fixture(n)) for n in names]


def golden_json(name):
    with open(os.path.join(GOLD, name)) as fh:
        return json.load(fh)


def random_cases():
    """(cases list, arrays) recorded from the unmodified reference by oracle/make_golden.py."""
    cases = golden_json("random_cases.json")
    arrays = np.load(os.path.join(GOLD, "random_cases.npz"))
    return cases, arrays


def case_inputs(case, arrays):
    """Packed GT (all rows, incl. uninformative), AF, names for one recorded random case."""
    cid = case["case"]
    packed = arrays[f"gt_{cid}"]
    af = arrays[f"af_{cid}"]
    names = np.array([f"S{i:05d}" for i in range(case["n_samples"])])
    return packed, af, names


def weights_table(path):
    """weights.txt of the reference as {name: weight} (utmos/select.py:343-352)."""
    out = {}
    with open(path) as fh:
        for line in fh:
            if line.strip():
                name, val = line.rstrip("\n").split("\t")
                out[name] = float(val)
    return out


# Hand-built VCF rows for the genotype forms the reference's fixtures do not contain, with the answers scikit-allel 1.3.5
# gives for them by its documented definitions (GenotypeArray.is_het: all alleles called and not all equal; is_hom_alt:
# all called, all equal, > 0; count_alleles counts every called allele, also of half-missing calls; to_frequencies =
# count / called; is_singleton(a) = count of allele a == 1) -- utmos/convert.py:57-77.  Samples A B C D.
HAND_VCF_HEAD = ("##fileformat=VCFv4.2\n##contig=<ID=1>\n"
                 "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tA\tB\tC\tD\n")
HAND_VCF_ROWS = [
    # line,                                              presence,      AF,        het, hom, singleton
    ("1\t10\t.\tA\tC\t.\t.\t.\tGT\t0|1\t1/1\t./.\t0|0", [1, 1, 0, 0], 3 / 6, 1, 1, False),   # missing call
    ("1\t11\t.\tA\tC\t.\t.\t.\tGT\t./1\t1|.\t0/1\t.", [0, 0, 1, 0], 3 / 4, 1, 0, True),      # half-missing: allele counted, call absent
    ("1\t12\t.\tA\tC\t.\t.\t.\tGT\t./.\t./.\t.\t./.", [0, 0, 0, 0], float("nan"), 0, 0, False),   # AN = 0 -> NaN
    ("1\t13\t.\tA\tC\t.\t.\t.\tGT\t1\t0\t1\t.", [0, 0, 0, 0], 2 / 3, 0, 0, True),            # haploid calls: second allele missing
    ("1\t14\t.\tA\tC,G,T\t.\t.\t.\tGT\t1|2\t2|2\t0|2\t3|3", [1, 1, 1, 1], 4 / 8, 2, 2, True),  # multi-allelic: max alt frequency
    ("1\t15\t.\tA\tC\t.\t.\t.\tGT\t0|0\t0|0\t0|0\t0|0", [0, 0, 0, 0], 0.0, 0, 0, False),     # all reference
    ("1\t16\t.\tA\tC\t.\t.\t.\tGT\t0|0\t0|1\t0|0\t0|0", [0, 1, 0, 0], 1 / 8, 1, 0, True),    # a plain singleton
]
