#!/usr/bin/env python3
"""How many steps of a --af selection does the reference-tie replay decide?  (counter 12 of utmos_debug_counters)"""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from utmos_b200 import _native, synth

n_samples, n_vars = 2504, 1_103_547
coh = synth.DeviceCohort(0, n_vars, n_samples)
names = synth.sample_names(n_samples)
weights = synth.synthetic_weights(n_samples)
mask = np.where(np.isin(names, names[: n_samples // 2]), 1, 2).astype(np.uint8)
mask = np.where(np.isin(names, names[::97]), 2, mask).astype(np.uint8)
out = {}
for label, m, w in (("c3_options", mask, weights), ("af_only", np.ones(n_samples, np.uint8), None)):
    dm = _native.DeviceMatrix(n_samples, _native.AF_F64, rows_hint=n_vars, flags=_native.F_REF_TIES)
    dm.append_packed_device(coh.rows.ptr, n_vars, coh.pitch, coh.af.ptr)
    dm.finalize()
    dm.begin(m, w)
    idx, new, score, stop = dm.steps(n_samples)
    cnt = dm.counters()
    tie_steps = int(cnt[12])
    out[label] = {"picks": int(len(idx)), "steps_decided_by_replay": tie_steps, "ref_ties": dm.info()["ref_ties"],
                  "first_tie_new_counts": None}
    print(label, out[label], flush=True)
    dm.close()
json.dump(out, open("gpurun_out/r2_tie_steps.json", "w"), indent=1)
