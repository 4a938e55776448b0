#!/usr/bin/env python3
"""Build container only (needs /root/reference): the UNMODIFIED reference's calculate_scores (utmos/select.py:24-53)
timed next to the NumPy port that bench.py uses as its CPU baseline (oracle.select_oracle.DenseOracle), on the same
dense slab of the synthetic cohort and the same greedy steps, plus one full `--count -1` run of the reference at a
reduced shape.  Shows that the port is a faithful stand-in where the reference cannot travel (the GPU box has no
/root/reference).  One JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refloader, select_oracle as orc  # noqa: E402  pylint: disable=wrong-import-position
from utmos_b200 import synth  # noqa: E402  pylint: disable=wrong-import-position


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=65536)
    ap.add_argument("--samples", type=int, default=2504)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--full-rows", type=int, default=20000, help="rows of the full --count -1 reference run (0 = skip)")
    args = ap.parse_args()
    ref = refloader.load_reference_select()
    gt, _ = synth.mirror_rows(0, 0, args.rows, args.samples)
    dense = np.unpackbits(gt, axis=1, count=args.samples).astype(bool)
    out = {"rows": args.rows, "samples": args.samples, "steps": args.steps, "host_cores": os.cpu_count()}
    # reference
    mask = np.ones(args.samples, dtype=np.uint8)
    weights = np.ones(args.samples)
    picks_ref, t_ref = [], []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        best, _new = ref.calculate_scores(dense, mask, weights)
        t_ref.append(time.perf_counter() - t0)
        picks_ref.append(int(best))
        mask[best] = 0
    # port
    oracle = orc.DenseOracle(dense, np.ones(args.samples, dtype=np.uint8))
    picks_port, t_port = [], []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        best = oracle.score_once()[0]
        t_port.append(time.perf_counter() - t0)
        picks_port.append(int(best))
        oracle.mask[best] = 0
    assert picks_ref == picks_port
    out["reference_s_per_step"] = t_ref
    out["port_s_per_step"] = t_port
    out["port_over_reference"] = float(np.mean(t_port) / np.mean(t_ref))
    if args.full_rows:
        gt, _ = synth.mirror_rows(0, 0, args.full_rows, args.samples)
        dense = np.unpackbits(gt, axis=1, count=args.samples).astype(bool)
        data = {"data": dense, "samples": synth.sample_names(args.samples).astype("S"), "var_count": dense.sum(axis=0)}
        t0 = time.perf_counter()
        rows = list(ref.run_selection(data, -1, None, None, None))
        out["full_run"] = {"rows": args.full_rows, "picks": len(rows), "seconds": time.perf_counter() - t0,
                           "last_row": [str(x) for x in rows[-1]]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
