#!/bin/bash
# 8 GPUs: list budget 2^26 (new default) against 2^24, and a later hand-over point
set -u
mkdir -p gpurun_out
N=8
run() {
  name=$1; shift
  timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 4 --warmup 3 --no-cpu "$@" > gpurun_out/r2_ho_$name.json 2> gpurun_out/r2_ho_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_ho_$name.json").read().strip().splitlines()[-1])
    print("$name", "ms", round(d["ms_per_step"],3), "verified", d.get("verified_vs_single_gpu"), d["select_parts_ms"], d["host_ms_create_append_finalize_begin_steps_close"]["resident"])
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r2_ho_$name.err").read()[-1200:])
PY
}
run n8_b26
run n8_b24 --list-budget 16777216 --no-verify
run n8_b26_r24k --tail-rows 24576 --no-verify
