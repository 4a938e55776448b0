# round 2, GPU call 15 (1 GPU): K2 flavour 4 (persistent ring of TMA stages): parity subset, then A/B against flavour 3
export UTMOS_B200_INGEST=4
timeout -k 10 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ingest_many or ragged or empty_and or synthetic_reduced or cli_answer or jl2" > gpurun_out/r2_pytest_ingest4.log 2>&1; echo "pytest(ingest 4) rc=$?"; tail -3 gpurun_out/r2_pytest_ingest4.log
: > gpurun_out/r2_streaming_ring.jsonl
for t in 20480 26624 34816 49152 69632; do UTMOS_B200_INGEST_RING_TILE=$t timeout -k 5 90 python tools/bench_streaming.py --tag ring_$t >> gpurun_out/r2_streaming_ring.jsonl; done
UTMOS_B200_INGEST_RING_TILE=34816 timeout -k 5 120 python tools/bench_streaming.py --tag ring_s100k --samples 100000 --vars 400000 --reps 3 >> gpurun_out/r2_streaming_ring.jsonl
UTMOS_B200_INGEST=3 python tools/bench_streaming.py --tag flavour3 >> gpurun_out/r2_streaming_ring.jsonl
python - <<'PY'
import json
for l in open("gpurun_out/r2_streaming_ring.jsonl"):
    try:
        d=json.loads(l); print(d["tag"], round(d["ingest"]["ms"],4), round(d["ingest"]["frac"],3))
    except Exception as e: print("bad line", l[:200])
PY
