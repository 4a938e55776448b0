#!/bin/bash
# AF flavour of the entry-divided tail with native 64-bit remote atomics: parity, threshold sweep on C3; C2 re-check
set -u
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tail_flavours or does_not_fit_one_sm or full_shape" > gpurun_out/r2_lc_pytest3.log 2>&1
echo "parity rc=$?"; tail -4 gpurun_out/r2_lc_pytest3.log
one() {
  name=$1; shift
  timeout -k 10 300 python bench.py --no-cpu --steps 4 "$@" > gpurun_out/r2_t9_$name.json 2> gpurun_out/r2_t9_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_t9_$name.json").read().strip().splitlines()[-1])
    print("$name", "ms", round(d["ms_per_step"],3), d["select_parts_ms"], d.get("verified_vs_oracle_golden"))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r2_t9_$name.err").read()[-1200:])
PY
}
one c2
one c2_h0 --heavy-rows 0
one c3_h0 --config c3 --heavy-rows 0
one c3_h512 --config c3 --heavy-rows 512
one c3_h1024 --config c3 --heavy-rows 1024
one c3_h1536 --config c3 --heavy-rows 1536
