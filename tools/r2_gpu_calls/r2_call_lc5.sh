#!/bin/bash
# earlier hand-over to the (now entry-divided) tail: tail_rows x list budget sweep on 2 GPUs and on one GPU with 8x rows
set -u
mkdir -p gpurun_out
N=2
run() {  # name, extra args
  name=$1; shift
  timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu "$@" > gpurun_out/r2_ho_$name.json 2> gpurun_out/r2_ho_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_ho_$name.json").read().strip().splitlines()[-1])
    print("$name", "ms", round(d["ms_per_step"],3), "verified", d.get("verified_vs_single_gpu"), d["select_parts_ms"])
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r2_ho_$name.err").read()[-1200:])
PY
}
run n2_default
run n2_rows8k --tail-rows 8192
run n2_rows32k --tail-rows 32768
run n2_rows32k_b26 --tail-rows 32768 --list-budget 67108864
run n2_rows128k_b26 --tail-rows 131072 --list-budget 67108864
run n2_rows128k_b27 --tail-rows 131072 --list-budget 134217728
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
one c2_rows8k --tail-rows 8192
one c2_rows32k_b26 --tail-rows 32768 --list-budget 67108864
one big_default --vars 8828376
one big_rows64k --vars 8828376 --tail-rows 65536
one big_rows64k_b26 --vars 8828376 --tail-rows 65536 --list-budget 67108864
one big_rows256k_b27 --vars 8828376 --tail-rows 262144 --list-budget 134217728
