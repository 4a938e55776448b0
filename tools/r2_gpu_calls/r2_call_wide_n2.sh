#!/bin/bash
# wide cohort (100,000 samples) with the entry-divided tail: 2 GPUs against 1 GPU, results must be identical
set -u
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout -k 10 400 python tools/run_configs.py c5 --vars 2000000 --out gpurun_out/r2_wide_n1.npz 2>&1 | tail -1
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29530 tools/run_configs.py c5 --vars 2000000 --out gpurun_out/r2_wide_n2.npz 2>&1 | tail -1
python tools/diff_runs.py gpurun_out/r2_wide_n1.npz gpurun_out/r2_wide_n2.npz
# AF multi-GPU (C3 options) through bench's verification leg
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --config c3 --steps 3 --warmup 2 --no-cpu > gpurun_out/r2_c3_n2.json 2> gpurun_out/r2_c3_n2.err; echo "c3 n2 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2_c3_n2.json").read().strip().splitlines()[-1])
    print("c3 n2 ms", round(d["ms_per_step"],3), "verified_vs_single_gpu", d.get("verified_vs_single_gpu"), d["select_parts_ms"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2_c3_n2.err").read()[-1500:])
PY
