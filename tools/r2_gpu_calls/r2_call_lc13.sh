#!/bin/bash
# AF (C3): earlier hand-over to the entry-divided tail
set -u
mkdir -p gpurun_out
one() {
  name=$1; shift
  timeout -k 10 300 python bench.py --no-cpu --steps 4 "$@" > gpurun_out/r2_t13_$name.json 2> gpurun_out/r2_t13_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_t13_$name.json").read().strip().splitlines()[-1])
    print("$name", "ms", round(d["ms_per_step"],3), d["select_parts_ms"], d.get("verified_vs_oracle_golden"))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r2_t13_$name.err").read()[-1200:])
PY
}
one c3_r4k --config c3 --tail-rows 4096
one c3_r8k --config c3 --tail-rows 8192
one c3_r16k --config c3 --tail-rows 16384
one c3_r64k --config c3 --tail-rows 65536
one c3_r1m --config c3 --tail-rows 1048576
