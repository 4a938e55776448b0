#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ref_ties" > gpurun_out/r2_reft_pytest.log 2>&1
echo "ref-ties tests rc=$?"; tail -25 gpurun_out/r2_reft_pytest.log
timeout -k 10 300 python bench.py --config c3 --no-cpu --steps 3 > gpurun_out/r2_reft_c3.json 2> gpurun_out/r2_reft_c3.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2_reft_c3.json").read().strip().splitlines()[-1])
    print("c3 exact ms", round(d["ms_per_step"],3), d["select_parts_ms"], d.get("verified_vs_oracle_golden"))
    print("ref order", d["configs"]["c3"]["reference_tie_order"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2_reft_c3.err").read()[-1500:])
PY
