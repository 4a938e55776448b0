#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tail_flavours or does_not_fit_one_sm or full_shape or list_budget" > gpurun_out/r2_lc_pytest5.log 2>&1
echo "parity rc=$?"; tail -4 gpurun_out/r2_lc_pytest5.log
one() {
  name=$1; shift
  timeout -k 10 300 python bench.py --no-cpu --steps 4 "$@" > gpurun_out/r2_t11_$name.json 2> gpurun_out/r2_t11_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_t11_$name.json").read().strip().splitlines()[-1])
    print("$name", "ms", round(d["ms_per_step"],3), d["select_parts_ms"], d.get("verified_vs_oracle_golden"))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r2_t11_$name.err").read()[-1200:])
PY
}
one c2_h1 --heavy-rows 1
one c2_h256 --heavy-rows 256
one c2
one c3_h1 --config c3 --heavy-rows 1
one c3 --config c3
one big --vars 8828376 --no-verify
