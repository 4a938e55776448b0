#!/bin/bash
# entry-divided cluster tail: parity first, then A/B on C2 and a large-V shape
set -u
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tail_flavours or does_not_fit_one_sm" > gpurun_out/r2_lc_pytest.log 2>&1
echo "parity rc=$?"; tail -8 gpurun_out/r2_lc_pytest.log
for h in 0 128 256 512 1024 2048; do
  timeout -k 10 200 python bench.py --no-cpu --steps 5 --heavy-rows $h > gpurun_out/r2_lc_c2_h$h.json 2>gpurun_out/r2_lc_c2_h$h.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_lc_c2_h$h.json").read().strip().splitlines()[-1])
    print("c2 heavy=$h", round(d["ms_per_step"],3), d["select_parts_ms"], d.get("verified_vs_oracle_golden"), d["gpu_launches"])
except Exception as e:
    print("c2 heavy=$h failed", e); print(open("gpurun_out/r2_lc_c2_h$h.err").read()[-800:])
PY
done
for h in 0 512; do
  timeout -k 10 300 python bench.py --no-cpu --steps 3 --config c3 --heavy-rows $h > gpurun_out/r2_lc_c3_h$h.json 2>gpurun_out/r2_lc_c3_h$h.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_lc_c3_h$h.json").read().strip().splitlines()[-1])
    print("c3 heavy=$h", round(d["ms_per_step"],3), d["select_parts_ms"], d.get("verified_vs_oracle_golden"))
except Exception as e:
    print("c3 heavy=$h failed", e); print(open("gpurun_out/r2_lc_c3_h$h.err").read()[-800:])
PY
done
for h in 0 512; do
  timeout -k 10 300 python bench.py --no-cpu --no-verify --steps 3 --vars 8828376 --heavy-rows $h > gpurun_out/r2_lc_big_h$h.json 2>gpurun_out/r2_lc_big_h$h.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_lc_big_h$h.json").read().strip().splitlines()[-1])
    print("8.8M heavy=$h", round(d["ms_per_step"],3), d["select_parts_ms"])
except Exception as e:
    print("8.8M heavy=$h failed", e); print(open("gpurun_out/r2_lc_big_h$h.err").read()[-800:])
PY
done
