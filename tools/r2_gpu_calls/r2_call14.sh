# round 2, GPU call 14 (1 GPU): AF decrement kernel for heavy picks, unaligned direct convert kernel
timeout -k 10 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_g.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_gpu_g.log
run() { tag=$1; shift; timeout -k 10 300 python bench.py --no-cpu --no-verify --steps 3 "$@" > gpurun_out/r2_b14_$tag.json 2> gpurun_out/r2_b14_$tag.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_b14_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["select_parts_ms"].items()}, d["run_config"]["greedy_steps"], d["verified_vs_oracle_golden"], d["gpu_launches"])
except Exception as e:
    print("$tag", "failed", e, open("gpurun_out/r2_b14_$tag.err").read()[-400:])
PY
}
run c3 --config c3
export UTMOS_B200_DECREMENT=0; run c3_regain --config c3; unset UTMOS_B200_DECREMENT
run c3_rg2048 --config c3 --regain-rows 2048
run c3_rg1024 --config c3 --regain-rows 1024
run c2
python tools/bench_convert.py --samples 2500 > gpurun_out/r2_convert_unaligned.json 2>&1; cut -c1-330 gpurun_out/r2_convert_unaligned.json
UTMOS_B200_CONVERT_TILE=1 python tools/bench_convert.py --samples 2500 > gpurun_out/r2_convert_tile_b.json 2>&1; cut -c1-330 gpurun_out/r2_convert_tile_b.json
python tools/bench_convert.py --samples 2501 > gpurun_out/r2_convert_unaligned_odd.json 2>&1; cut -c1-330 gpurun_out/r2_convert_unaligned_odd.json
python tools/bench_convert.py > gpurun_out/r2_convert_b.json 2>&1; cut -c1-330 gpurun_out/r2_convert_b.json
