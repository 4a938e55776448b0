# round 2, final single-GPU call: parity suite, smoke, bench lines, ncu launch list and full capture of the tail kernel
timeout -k 10 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke.log
timeout -k 10 400 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"
timeout -k 10 300 python bench.py --impl reference > gpurun_out/r2_bench_reference_arm.json 2>&1; echo "reference arm rc=$?"
timeout -k 10 300 python bench.py --pageable --no-cpu --no-verify > gpurun_out/r2_bench_final_pageable.json 2>/dev/null
timeout -k 10 300 python bench.py --config c3 --no-cpu --no-verify > gpurun_out/r2_bench_final_c3.json 2>/dev/null
timeout -k 10 300 python bench.py --step-times --no-cpu --no-verify --steps 2 > gpurun_out/r2_bench_final_times.json 2>/dev/null; cp gpurun_out/step_series.json gpurun_out/r2_step_series_final.json
python bench.py --steps 2 --warmup 1 --no-cpu --no-verify > gpurun_out/r2_plain_for_ncu.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-verify > gpurun_out/r2_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python bench.py --steps 1 --warmup 1 --no-cpu --no-verify > gpurun_out/r2_plain_for_ncu2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:select_tail_kernel -s 6 -c 6 -o gpurun_out/r2_tail_final python bench.py --steps 1 --warmup 1 --no-cpu --no-verify > gpurun_out/r2_ncu_tail.log 2>&1
echo "ncu tail rc=$?"
python bench.py --config c3 --steps 1 --warmup 1 --no-cpu --no-verify > gpurun_out/r2_plain_for_ncu3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:select_listcluster_kernel -s 1 -c 1 -o gpurun_out/r2_entry_cluster_c3 python bench.py --config c3 --steps 1 --warmup 1 --no-cpu --no-verify > gpurun_out/r2_ncu_entry_cluster.log 2>&1
echo "ncu entry-divided tail rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_final.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], d["select_parts_ms"], "golden", d["verified_vs_oracle_golden"], "launches", d["gpu_launches"])
print({k:(round(v["frac"],3), round(v["ms"],4)) for k,v in d["streaming_kernels"].items()}, d["roofline"]["frac"], d["clocks"])
print("c3", d["configs"]["c3"]["ms_per_selection"], d["configs"]["c3"]["us_per_greedy_step"], d["configs"]["c3"]["verified_vs_oracle_golden"])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["like_for_like"])
r=json.loads(open("gpurun_out/r2_bench_reference_arm.json").read().strip().splitlines()[-1])
print("ref arm", r["value"], r["ms_per_step"], r["cpu_baseline"]["kind"])
PY
