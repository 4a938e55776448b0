#!/bin/bash
set -u
mkdir -p gpurun_out
for h in 1 128; do
  timeout -k 10 200 python bench.py --no-cpu --steps 3 --heavy-rows $h > gpurun_out/r2_lcp_c2_h$h.json 2>gpurun_out/r2_lcp_c2_h$h.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_lcp_c2_h$h.json").read().strip().splitlines()[-1])
    print("c2 heavy=$h", round(d["ms_per_step"],3), d["select_parts_ms"], d.get("verified_vs_oracle_golden"), d["phase_cycles"])
except Exception as e:
    print("c2 heavy=$h failed", e); print(open("gpurun_out/r2_lcp_c2_h$h.err").read()[-800:])
PY
done
