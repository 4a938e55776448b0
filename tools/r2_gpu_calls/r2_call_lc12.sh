#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_entry_cluster.log 2>&1
echo "full suite rc=$?"; tail -4 gpurun_out/r2_pytest_gpu_entry_cluster.log
one() {
  name=$1; shift
  timeout -k 10 300 python bench.py --no-cpu --steps 4 "$@" > gpurun_out/r2_t12_$name.json 2> gpurun_out/r2_t12_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_t12_$name.json").read().strip().splitlines()[-1])
    print("$name", "ms", round(d["ms_per_step"],3), d["select_parts_ms"], d.get("verified_vs_oracle_golden"))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r2_t12_$name.err").read()[-1200:])
PY
}
one c2
one c3 --config c3
one c3_h64 --config c3 --heavy-rows 64
