#!/usr/bin/env python3
"""K2 / K2b / K3 alone (ingest, transpose, column reduce) on the synthetic cohort already resident in HBM: CUDA-event
times of the library's own launches (utmos_timings), best of --reps, against the measured HBM peak.  Kernel flavours
are chosen by the environment (UTMOS_B200_INGEST, UTMOS_B200_INGEST_TILE, UTMOS_B200_TRANSPOSE), one process each."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from utmos_b200 import _native, synth  # noqa: E402  pylint: disable=wrong-import-position


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vars", type=int, default=1_103_547)
    ap.add_argument("--samples", type=int, default=2504)
    ap.add_argument("--reps", type=int, default=6)
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    cohort = synth.DeviceCohort(0, args.vars, args.samples)
    pitch = (args.samples + 7) // 8
    best = None
    var_count = None
    for _ in range(args.reps):
        dm = _native.DeviceMatrix(args.samples, _native.AF_NONE, rows_hint=args.vars)
        dm.append_packed_device(cohort.rows.ptr, args.vars, pitch)
        vc = dm.finalize()
        ms = dm.timings()
        assert var_count is None or (vc == var_count).all()
        var_count = vc
        best = ms if best is None else {k: min(best[k], ms[k]) for k in ms}
        rows = dm.num_vars
        dm.close()
    peak = 6548.2
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except OSError:
        pass
    pitch_out = (args.samples + 127) // 128 * 16
    alg = {"ingest": args.vars * pitch + rows * pitch_out, "transpose": 2 * rows * pitch_out, "colreduce": rows * pitch_out}
    out = {"tag": args.tag, "env": {k: v for k, v in os.environ.items() if k.startswith("UTMOS_B200_")}, "rows": int(rows),
           "var_count_sum": int(var_count.sum())}
    for name, key in (("ingest", "ingest_ms"), ("transpose", "transpose_ms"), ("colreduce", "gain_ms")):
        out[name] = {"ms": best[key], "bytes": alg[name], "GBps": alg[name] / 1e6 / best[key], "frac": alg[name] / 1e6 / best[key] / peak}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
