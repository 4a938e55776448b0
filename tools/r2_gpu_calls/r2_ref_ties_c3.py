#!/usr/bin/env python3
"""Config C3 (2,504 x 1,103,547, --af --weights --subset --exclude, --count -1) with UTMOS_F_REF_TIES: time and how far the
reference-order replay moves the ordering away from the exact-arithmetic golden (tests/golden/c3_full_order.npz)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from utmos_b200 import _native, synth  # noqa: E402
import bench  # noqa: E402

n_vars, n_samples = 1_103_547, 2504
coh = synth.DeviceCohort(0, n_vars, n_samples)
mask, weights = bench.c3_options(n_samples)
gold = np.load(os.path.join(ROOT, "tests", "golden", "c3_full_order.npz"))
out = {}
for label, flags in (("exact_default", 0), ("ref_ties", _native.F_REF_TIES), ("step_kernels_exact", _native.F_STEP_KERNELS)):
    dm = _native.DeviceMatrix(n_samples, _native.AF_F64, rows_hint=n_vars, flags=flags)
    dm.append_packed_device(coh.rows.ptr, n_vars, coh.pitch, coh.af.ptr)
    dm.finalize()
    _native.timer_start(0)
    dm.begin(mask, weights)
    idx, new, score, stop = dm.steps(n_samples)
    ms = _native.timer_stop(0)
    same = int(np.sum(idx[:len(gold["idx"])] == gold["idx"][:len(idx)]))
    first_diff = int(np.argmax(idx[:len(gold["idx"])] != gold["idx"][:len(idx)])) if same < len(idx) else -1
    out[label] = {"select_ms": ms, "picks": int(len(idx)), "stop": int(stop), "same_positions_as_exact_golden": same,
                  "first_difference_at_pick": first_diff,
                  "max_rel_score_diff": float(np.max(np.abs(score[:same] - gold["score"][:same]) / gold["score"][:same])) if same else None}
    dm.close()
print(json.dumps(out))
