#!/usr/bin/env python3
"""Generate tests/golden/* by running the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE.  Run once here (``python oracle/make_golden.py``); the outputs are committed because
/root/reference does not exist on the GPU box.

Produces
  tests/golden/ref/test_files/*   the reference's own fixtures (data files, copied byte for byte)
  tests/golden/ref/answer_key/*   the reference's own answer keys
  tests/golden/full_order_count.json / full_order_af.json
        ``--count -1`` orderings of chunk0.jl+chunk1.jl from the reference, with the winning score of every
        step captured by wrapping ``numpy.argmax`` (called at utmos/select.py:48)
  tests/golden/random_cases.npz + random_cases.json
        seeded random matrices (ties, exclusions, weights, --af, float32-af) and what the reference answers
"""
import io
import json
import os
import shutil
import sys
import tempfile

import joblib
import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle.refloader import REFERENCE_ROOT, load_reference_select  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
FIXTURES = ["chunk0.jl", "chunk1.jl", "chunk2.jl", "chunk0.vcf.gz", "chunk1.vcf.gz", "chunk_tiny.vcf",
            "weights.txt", "subset.txt", "exclude.txt", "tiny.hdf5", "tiny.af.hdf5"]


def copy_fixtures():
    """Reference data files + answer keys, byte for byte."""
    dst = os.path.join(GOLD, "ref", "test_files")
    os.makedirs(dst, exist_ok=True)
    for name in FIXTURES:
        shutil.copyfile(os.path.join(REFERENCE_ROOT, "repo_utils", "test_files", name), os.path.join(dst, name))
    dst = os.path.join(GOLD, "ref", "answer_key")
    os.makedirs(dst, exist_ok=True)
    src = os.path.join(REFERENCE_ROOT, "repo_utils", "answer_key")
    for name in sorted(os.listdir(src)):
        shutil.copyfile(os.path.join(src, name), os.path.join(dst, name))


class ArgmaxSpy:
    """Records scores[argmax] for every np.argmax call made while active (utmos/select.py:48)."""

    def __init__(self):
        self.best_scores = []
        self._orig = None

    def __enter__(self):
        self._orig = np.argmax

        def spy(arr, *args, **kwargs):
            out = self._orig(arr, *args, **kwargs)
            self.best_scores.append(float(np.asarray(arr)[out]))
            return out
        np.argmax = spy
        return self

    def __exit__(self, *exc):
        np.argmax = self._orig


def run_reference(ref, files, count, af=False, subset=None, exclude=None, weights_file=None):
    """load_files + run_selection of the reference; returns (rows, best score per argmax call)."""
    data = ref.load_files(files, None, 32768, af)
    weights = None
    if weights_file:
        # select.py:187 assigns a 1-row DataFrame slice into a scalar slot, which numpy>=2 rejects;
        # handing the Series over performs the same lookup without editing the reference.
        weights = ref.parse_weights(weights_file)["weight"]
    with ArgmaxSpy() as spy:
        rows = list(ref.run_selection(data, count, subset, exclude, weights))
    return rows, spy.best_scores


def rows_to_json(rows):
    return [[str(r[0]), int(r[1]), int(r[2]), int(r[3]), float(r[4]), str(r[4])] for r in rows]


def full_orderings(ref):
    tf = os.path.join(REFERENCE_ROOT, "repo_utils", "test_files")
    files = [os.path.join(tf, "chunk0.jl"), os.path.join(tf, "chunk1.jl")]
    for name, af in (("full_order_count", False), ("full_order_af", True)):
        rows, scores = run_reference(ref, files, -1, af=af)
        with open(os.path.join(GOLD, name + ".json"), "w") as fh:
            json.dump({"files": ["chunk0.jl", "chunk1.jl"], "count": -1, "af": af,
                       "rows": rows_to_json(rows), "argmax_scores": scores}, fh)
        print(name, len(rows), "rows")


def random_matrix(rng, n_vars, n_samples, style):
    """Bool matrices that exercise ties: private singletons, duplicated columns, empty rows."""
    if style == "sparse":
        dens = rng.uniform(0.01, 0.2)
        mat = rng.random((n_vars, n_samples)) < dens
    elif style == "powerlaw":
        k = np.clip((rng.pareto(0.7, n_vars) + 1).astype(int), 1, n_samples)
        mat = np.zeros((n_vars, n_samples), dtype=bool)
        for v in range(n_vars):
            mat[v, rng.choice(n_samples, k[v], replace=False)] = True
    elif style == "ties":
        mat = np.zeros((n_vars, n_samples), dtype=bool)
        mat[np.arange(n_vars), rng.integers(0, n_samples, n_vars)] = True      # private singletons
        extra = rng.random((n_vars, n_samples)) < 0.02
        mat |= extra
        if n_samples > 3:
            mat[:, n_samples // 2] = mat[:, 1]                                   # duplicate column
    else:
        raise ValueError(style)
    drop = rng.random(n_vars) < 0.05
    mat[drop] = False                                                            # uninformative rows
    return mat


def random_cases(ref):
    rng = np.random.default_rng(20261018)
    cases = []
    arrays = {}
    tmp = tempfile.mkdtemp()
    shapes = [(40, 5), (97, 36), (200, 33), (300, 64), (257, 130), (500, 257), (64, 8), (10, 1), (120, 100),
              (333, 31), (50, 129), (600, 96)]
    styles = ["sparse", "powerlaw", "ties"]
    case_id = 0
    for n_vars, n_samples in shapes:
        for style in styles:
            mat = random_matrix(rng, n_vars, n_samples, style)
            names = np.array([f"S{i:05d}" for i in range(n_samples)])
            # allele frequency of an imaginary diploid cohort: k/(2S) with repeated values -> exact ties
            carriers = mat.sum(axis=1)
            af = np.where(carriers > 0, np.maximum(carriers, 1) * rng.integers(1, 3, n_vars) / (2.0 * n_samples), 0.0)
            af = np.minimum(af, 1.0)
            if style == "ties":
                zero_af = rng.random(n_vars) < 0.03
                af[zero_af] = 0.0           # old-AF-definition rows: informative, AF == 0
            nfiles = 1 if n_vars < 100 else 2
            cuts = [0, n_vars] if nfiles == 1 else [0, n_vars // 3, n_vars]
            files = []
            for fi in range(nfiles):
                sl = slice(cuts[fi], cuts[fi + 1])
                path = os.path.join(tmp, f"case{case_id}_{fi}.jl")
                joblib.dump({"GT": np.packbits(mat[sl], axis=1), "samples": names,
                             "AF": af[sl].reshape(-1, 1), "stats": {}}, path)
                files.append(path)
            arrays[f"gt_{case_id}"] = np.packbits(mat, axis=1)
            arrays[f"af_{case_id}"] = af
            arrays[f"cuts_{case_id}"] = np.array(cuts)
            variants = [
                dict(count=-1, af=False, subset=None, exclude=None, weights=None),
                dict(count=-1, af=True, subset=None, exclude=None, weights=None),
                dict(count=0.5, af=False, subset=None, exclude=[names[0], names[n_samples // 2]], weights=None),
                dict(count=-1, af=True, subset=[str(x) for x in names[::2]], exclude=[names[0]],
                     weights={str(names[n_samples // 3]): 4, str(names[-1]): 10, "NOPE": 3}),
                dict(count=-1, af=False, subset=None, exclude=None,
                     weights={str(names[n_samples // 3]): 2.5, str(names[-1]): 0.0, str(names[0]): 7}),
            ]
            for var in variants:
                wfile = None
                if var["weights"] is not None:
                    wfile = os.path.join(tmp, f"w{case_id}.txt")
                    pd.DataFrame(list(var["weights"].items())).to_csv(wfile, sep="\t", header=False, index=False)
                rows, scores = run_reference(ref, files, var["count"], af=var["af"], subset=var["subset"],
                                             exclude=var["exclude"], weights_file=wfile)
                cases.append({"case": case_id, "n_vars": n_vars, "n_samples": n_samples, "style": style,
                              "options": {k: (v if not isinstance(v, np.ndarray) else v.tolist())
                                          for k, v in var.items()},
                              "rows": rows_to_json(rows), "argmax_scores": scores})
            case_id += 1
    np.savez_compressed(os.path.join(GOLD, "random_cases.npz"), **arrays)
    with open(os.path.join(GOLD, "random_cases.json"), "w") as fh:
        json.dump(cases, fh)
    print("random cases:", len(cases))
    shutil.rmtree(tmp)


def main():
    import logging
    logging.disable(logging.CRITICAL)
    os.makedirs(GOLD, exist_ok=True)
    ref = load_reference_select()
    copy_fixtures()
    full_orderings(ref)
    random_cases(ref)


if __name__ == "__main__":
    main()
