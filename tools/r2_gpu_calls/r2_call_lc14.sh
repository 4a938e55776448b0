#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tail_flavours or full_shape" > gpurun_out/r2_lc_pytest6.log 2>&1
echo "parity rc=$?"; tail -3 gpurun_out/r2_lc_pytest6.log
one() {
  name=$1; shift
  timeout -k 10 300 python bench.py --no-cpu --steps 4 "$@" > gpurun_out/r2_t14_$name.json 2> gpurun_out/r2_t14_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_t14_$name.json").read().strip().splitlines()[-1])
    print("$name", "ms", round(d["ms_per_step"],3), d["select_parts_ms"], d.get("verified_vs_oracle_golden"))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r2_t14_$name.err").read()[-1200:])
PY
}
one c3 --config c3
one c3b --config c3
