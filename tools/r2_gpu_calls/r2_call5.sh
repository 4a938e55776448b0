# round 2, GPU call 5: parity suite, K2 flavours, K1, C3 tail flavours
timeout -k 10 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_c.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_gpu_c.log
: > gpurun_out/r2_streaming.jsonl
python tools/bench_streaming.py --tag ingest1 >> gpurun_out/r2_streaming.jsonl
UTMOS_B200_INGEST=2 UTMOS_B200_INGEST_TILE=55296 python tools/bench_streaming.py --tag ingest2_54k >> gpurun_out/r2_streaming.jsonl
for t in 27648 36864 55296; do UTMOS_B200_INGEST=3 UTMOS_B200_INGEST_TILE=$t timeout -k 5 60 python tools/bench_streaming.py --tag ingest3_$t >> gpurun_out/r2_streaming.jsonl; done
UTMOS_B200_INGEST=3 timeout -k 5 120 python tools/bench_streaming.py --tag ingest3_s100k --samples 100000 --vars 400000 --reps 3 >> gpurun_out/r2_streaming.jsonl
python - <<'PY'
import json
for l in open("gpurun_out/r2_streaming.jsonl"):
    try:
        d=json.loads(l); print(d["tag"], d["ingest"]["ms"], round(d["ingest"]["frac"],3))
    except Exception as e: print("bad line", l[:200])
PY
UTMOS_B200_INGEST=3 timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ingest_many or ragged or empty_and or synthetic_reduced or cli_answer" > gpurun_out/r2_pytest_ingest3.log 2>&1; echo "pytest(ingest 3) rc=$?"; tail -2 gpurun_out/r2_pytest_ingest3.log
python tools/bench_convert.py > gpurun_out/r2_convert_a.json 2>&1; cut -c1-330 gpurun_out/r2_convert_a.json
python tools/bench_convert.py --scatter > gpurun_out/r2_convert_scatter.json 2>&1; cut -c1-330 gpurun_out/r2_convert_scatter.json
python tools/bench_convert.py --samples 2500 > gpurun_out/r2_convert_tile.json 2>&1; cut -c1-330 gpurun_out/r2_convert_tile.json
run() { tag=$1; shift; timeout -k 10 150 python bench.py --no-cpu --no-verify --steps 3 "$@" > gpurun_out/r2_b5_$tag.json 2> gpurun_out/r2_b5_$tag.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_b5_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["select_parts_ms"].items()}, "gain_ms", round(d["phases_ms"]["gain_ms"],3), d["phase_cycles"][5:8], d["phase_cycles"][12], "golden", d["verified_vs_oracle_golden"])
except Exception as e:
    print("$tag", "failed", e, open("gpurun_out/r2_b5_$tag.err").read()[-300:])
PY
}
run c3 --config c3
run c3_sr64 --config c3 --single-rows 64
run c3_sr256 --config c3 --single-rows 256
run c3_sr1024 --config c3 --single-rows 1024
run c3_sr256_tr4096 --config c3 --single-rows 256 --tail-rows 4096
run c3_sr256_tr8192 --config c3 --single-rows 256 --tail-rows 8192
run c2
