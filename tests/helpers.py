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
