# round 2, GPU call 11 (1 GPU): full parity suite on the current defaults + bench lines (pinned, pageable)
timeout -k 10 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_d.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest_gpu_d.log
timeout -k 10 300 python bench.py > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2_bench_d.err
timeout -k 10 200 python bench.py --pageable --no-cpu --no-verify > gpurun_out/r2_bench_d_pageable.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_d.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], d["select_parts_ms"], d["streaming_kernels"])
print("c3", {k:d["configs"]["c3"][k] for k in ("ms_per_selection","us_per_greedy_step","verified_vs_oracle_golden","step0_gains")})
print("cpu", d["cpu_baseline"]["like_for_like"], d["verified_vs_oracle_golden"])
p=json.loads(open("gpurun_out/r2_bench_d_pageable.json").read().strip().splitlines()[-1])
print("pageable e2e ms", p["e2e"]["ms_per_step"], p["e2e"]["phases_ms"]["h2d_ms"])
PY
