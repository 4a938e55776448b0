#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -k 10 300 python bench.py --no-cpu --steps 3 > gpurun_out/r2_reft_c2.json 2> gpurun_out/r2_reft_c2.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2_reft_c2.json").read().strip().splitlines()[-1])
    print("c2 ms", round(d["ms_per_step"],3), d["select_parts_ms"], d.get("verified_vs_oracle_golden"))
    c3=d["configs"]["c3"]
    print("c3 exact", c3["ms_per_selection"], c3["verified_vs_oracle_golden"])
    print("ref order", c3["reference_tie_order"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2_reft_c2.err").read()[-1500:])
PY
