#!/usr/bin/env python3
"""Wide cohorts (state sliced over a cluster): owner-computes tail against the entry-divided tail.  A/B run, one GPU."""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from utmos_b200 import _native, synth

out = []
for n_samples, n_vars, count in ((50_000, 2_000_000, 20_000), (100_000, 2_000_000, 30_000)):
    coh = synth.DeviceCohort(0, n_vars, n_samples)
    ref = None
    for heavy in (0, -1, 1):
        dm = _native.DeviceMatrix(n_samples, _native.AF_NONE, rows_hint=n_vars)
        dm.append_packed_device(coh.rows.ptr, n_vars, coh.pitch, 0)
        dm.finalize()
        dm.set_option(10, heavy)
        mask = np.ones(n_samples, np.uint8)
        _native.timer_start(0)
        dm.begin(mask)
        idx, new, score, stop = dm.steps(count)
        ms = _native.timer_stop(0)
        tim = dm.timings()
        same = None
        if ref is None:
            ref = (idx, new, score, stop)
        else:
            same = bool(np.array_equal(idx, ref[0]) and np.array_equal(new, ref[1]) and np.array_equal(score, ref[2]) and stop == ref[3])
        rec = {"samples": n_samples, "rows": n_vars, "picks": int(len(idx)), "heavy_rows": heavy, "select_ms": round(ms, 2),
               "us_per_step": round(ms * 1e3 / max(len(idx), 1), 2), "same_as_heavy0": same, "flavour": dm.info()["flavour"],
               "parts_ms": {k: round(float(v), 2) for k, v in tim.items() if k in ("head_ms", "handover_ms", "tail_ms")}}
        print(json.dumps(rec), flush=True)
        out.append(rec)
        dm.close()
    coh.close()
json.dump(out, open("gpurun_out/r2_wide_ab.json", "w"), indent=1)
