#!/bin/bash
# entry-divided tail on 8 GPUs: bench at heavy-rows 1024 (default) and 512
set -u
mkdir -p gpurun_out
N=8
for h in 1024 512; do
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 5 --warmup 3 --heavy-rows $h > gpurun_out/r2_lc_n${N}_h$h.json 2> gpurun_out/r2_lc_n${N}_h$h.err; echo "bench n$N h$h rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_lc_n${N}_h$h.json").read().strip().splitlines()[-1])
    print(d["n_gpus"], "heavy $h value", round(d["value"],2), "ms", round(d["ms_per_step"],3), "verified", d["verified_vs_single_gpu"], d["select_parts_ms"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2_lc_n${N}_h$h.err").read()[-1500:])
PY
done
