#!/usr/bin/env python3
"""Full-shape golden orderings for configs C2 and C3 (BASELINE.json), from the plain-C oracle.  TEST INFRASTRUCTURE.

    python oracle/make_golden_full.py cohort      # seed-0 synthetic cohort 2,504 x 1,103,547 -> /tmp cache (NumPy mirror)
    python oracle/make_golden_full.py c2          # count mode, --count -1                -> tests/golden/c2_full_order.npz
    python oracle/make_golden_full.py c3          # --af --weights --subset --exclude -1  -> tests/golden/c3_full_order.npz
    python oracle/make_golden_full.py c3ref       # the same in the REFERENCE's float64 order -> tests/golden/c3ref_full_order.npz

The cohort is the NumPy mirror (utmos_b200/synth.py:mirror_rows) of the device generator, i.e. the rows bench.py and
tools/run_configs.py select from.  The orderings come from oracle/greedy_oracle.c (utmos/select.py:24-53, :69-112 restated
on packed bits; pinned to the unmodified reference by tests/test_oracle.py), one thread, minutes per config in the
build container.  C3 uses the exact (fixed-point) score mode, the arithmetic the CUDA --af path is bit-equal to; c3ref uses
mode 0 -- sequential float64 sums in row order, utmos/select.py:37-41, bit-identical to the reference's doubles -- which is
what `--ref-ties` (UTMOS_F_REF_TIES) reproduces: at this shape the two orders part at pick 225 of 1,239.
"""
import hashlib
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

N_SAMPLES, N_VARS, SEED = 2504, 1_103_547, 0
CACHE = os.environ.get("UTMOS_GOLDEN_CACHE", "/tmp/utmos_c2_seed0")
GOLD = os.path.join(ROOT, "tests", "golden")
BLOCK = 16384


def _block(r0):
    from utmos_b200 import synth
    n = min(BLOCK, N_VARS - r0)
    gt, af = synth.mirror_rows(SEED, r0, n, N_SAMPLES)
    return r0, gt, af


def make_cohort():
    pitch = (N_SAMPLES + 7) // 8
    rows = np.zeros((N_VARS, pitch), dtype=np.uint8)
    af = np.zeros(N_VARS, dtype=np.float64)
    with Pool(int(os.environ.get("UTMOS_GOLDEN_PROCS", "6"))) as pool:
        for r0, gt, a in pool.imap_unordered(_block, range(0, N_VARS, BLOCK)):
            rows[r0:r0 + len(gt)] = gt
            af[r0:r0 + len(gt)] = a
    np.save(CACHE + "_rows.npy", rows)
    np.save(CACHE + "_af.npy", af)
    print("cohort", rows.shape, hashlib.sha256(rows.tobytes()).hexdigest()[:16])


def c3_setup():
    """The C3 options exactly as tools/run_configs.py / bench.py --config c3 build them."""
    from utmos_b200 import synth
    names = synth.sample_names(N_SAMPLES)
    weights = synth.synthetic_weights(N_SAMPLES)
    mask = np.where(np.isin(names, names[: N_SAMPLES // 2]), 1, 2).astype(np.uint8)
    mask = np.where(np.isin(names, names[::97]), 2, mask).astype(np.uint8)
    return mask, weights


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def run(config):
    from oracle import select_oracle as orc
    orc.build_c_oracle()
    rows = np.load(CACHE + "_rows.npy", mmap_mode="r")
    rows = np.ascontiguousarray(rows)
    af = np.load(CACHE + "_af.npy")
    t0 = time.perf_counter()
    keep, var_count = orc.filter_rows_c(rows, N_SAMPLES)
    assert keep.all()                                      # every synthetic row is informative
    if config == "c2":
        mask, weights, afs, exact = np.ones(N_SAMPLES, np.uint8), None, None, False
    else:
        mask, weights = c3_setup()
        afs, exact = af, config == "c3"
    idx, new, score, stop = orc.greedy_c(rows, N_SAMPLES, mask, weights, afs, N_SAMPLES, exact=exact)
    sec = time.perf_counter() - t0
    out = os.path.join(GOLD, f"{config}_full_order.npz")
    np.savez_compressed(out, idx=idx.astype(np.int32), new=new.astype(np.int32), score=score, stop=np.int64(stop),
                        var_count=var_count.astype(np.int32), seed=np.int64(SEED), n_vars=np.int64(N_VARS),
                        n_samples=np.int64(N_SAMPLES), oracle_seconds=np.float64(sec),
                        sha256=np.array(digest(idx.astype(np.int64), new.astype(np.int64))))
    print(config, "steps", len(idx), "stop", stop, "tot", int(new.sum()), f"{sec:.1f} s", "->", out)


if __name__ == "__main__":
    what = sys.argv[1]
    if what == "cohort":
        make_cohort()
    else:
        run(what)
