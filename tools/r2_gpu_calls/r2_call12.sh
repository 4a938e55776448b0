# round 2, GPU call 12 (1 GPU): the tail at the size the replicated multi-GPU tail sees at N = 8 (8 x 1,103,547 rows on ONE GPU)
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "jl2 or pack2 or resume or export_import" > gpurun_out/r2_pytest_e.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_e.log
run() { tag=$1; shift; timeout -k 10 300 python bench.py --vars 8828376 --no-cpu --no-verify --steps 2 --warmup 1 "$@" > gpurun_out/r2_b12_$tag.json 2> gpurun_out/r2_b12_$tag.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_b12_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["select_parts_ms"].items()}, d["phase_cycles"][5:13], d["run_config"]["greedy_steps"])
except Exception as e:
    print("$tag", "failed", e, open("gpurun_out/r2_b12_$tag.err").read()[-400:])
PY
}
run default
run sr4096 --single-rows 4096
run tr4096 --tail-rows 4096
run tr16384 --tail-rows 16384
timeout -k 10 300 python bench.py --vars 8828376 --no-cpu --no-verify --steps 1 --warmup 1 --step-times > gpurun_out/r2_b12_times.json 2>/dev/null; cp gpurun_out/step_series.json gpurun_out/r2_step_series_8x.json
python bench.py --vars 8828376 --no-cpu --no-verify --steps 1 --warmup 0 > gpurun_out/r2_ncu12_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:select_tail_kernel -c 1 -o gpurun_out/r2_tail_8x python bench.py --vars 8828376 --no-cpu --no-verify --steps 1 --warmup 0 > gpurun_out/r2_ncu12.log 2>&1
echo "ncu rc=$?"
