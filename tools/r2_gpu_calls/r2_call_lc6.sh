#!/bin/bash
# list budget x hand-over rows on one GPU with 8x and 4x the rows
set -u
mkdir -p gpurun_out
one() {
  name=$1; shift
  timeout -k 10 300 python bench.py --no-cpu --no-verify --steps 3 "$@" > gpurun_out/r2_ho_$name.json 2> gpurun_out/r2_ho_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_ho_$name.json").read().strip().splitlines()[-1])
    print("$name", "ms", round(d["ms_per_step"],3), d["select_parts_ms"])
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r2_ho_$name.err").read()[-1200:])
PY
}
one big_b25 --vars 8828376 --list-budget 33554432
one big_b26 --vars 8828376 --list-budget 67108864
one big_b26_r32k --vars 8828376 --list-budget 67108864 --tail-rows 32768
one big_b26_r24k --vars 8828376 --list-budget 67108864 --tail-rows 24576
one big_b27_r16k --vars 8828376 --list-budget 134217728 --tail-rows 16384
one mid_default --vars 4414188
one mid_b25 --vars 4414188 --list-budget 33554432
one mid_b26 --vars 4414188 --list-budget 67108864
one c2_b25 --list-budget 33554432
