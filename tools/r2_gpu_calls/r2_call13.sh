# round 2, GPU call 13 (1 GPU): 32-bit shared-memory atomics for the AF limbs, RED for live bits in global memory
timeout -k 10 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_f.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_gpu_f.log
run() { tag=$1; shift; timeout -k 10 300 python bench.py --no-cpu --no-verify --steps 3 "$@" > gpurun_out/r2_b13_$tag.json 2> gpurun_out/r2_b13_$tag.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_b13_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["select_parts_ms"].items()}, d["run_config"]["greedy_steps"], d["verified_vs_oracle_golden"])
except Exception as e:
    print("$tag", "failed", e, open("gpurun_out/r2_b13_$tag.err").read()[-400:])
PY
}
run c2
run c3 --config c3
run c3_sr0 --config c3 --single-rows 0
run c2_8x --vars 8828376 --warmup 1 --steps 2
