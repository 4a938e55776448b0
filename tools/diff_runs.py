#!/usr/bin/env python3
"""Compare two runs saved by tools/run_configs.py --out (e.g. the 1-GPU and the N-GPU run of the same cohort):
report rows (sample index, new_count), winning scores, stop reason and var_count must be identical."""
import sys

import numpy as np

a, b = np.load(sys.argv[1]), np.load(sys.argv[2])
ok = True
for key in ("idx", "new", "score", "stop", "var_count"):
    same = np.array_equal(a[key], b[key])
    ok &= same
    print(f"{key:10s} {'identical' if same else 'DIFFERENT'}  shape {a[key].shape} vs {b[key].shape}")
print("steps", len(a["idx"]), "tot_captured", int(a["new"].sum()))
sys.exit(0 if ok else 1)
