python -m pytest tests -m gpu -x -q > gpurun_out/s40_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/s40_pytest.log
python __graft_entry__.py --smoke 2>&1 | tail -1
: > gpurun_out/s40_streaming.jsonl
python tools/bench_streaming.py --tag default >> gpurun_out/s40_streaming.jsonl 2>gpurun_out/s40_streaming.err
UTMOS_B200_TRANSPOSE=1 python tools/bench_streaming.py --tag tr1 >> gpurun_out/s40_streaming.jsonl 2>>gpurun_out/s40_streaming.err
UTMOS_B200_TRANSPOSE=0 UTMOS_B200_INGEST=0 python tools/bench_streaming.py --tag old >> gpurun_out/s40_streaming.jsonl 2>>gpurun_out/s40_streaming.err
UTMOS_B200_INGEST_TILE=13824 python tools/bench_streaming.py --tag tile13k >> gpurun_out/s40_streaming.jsonl 2>>gpurun_out/s40_streaming.err
UTMOS_B200_INGEST_TILE=55296 python tools/bench_streaming.py --tag tile54k >> gpurun_out/s40_streaming.jsonl 2>>gpurun_out/s40_streaming.err
python tools/bench_streaming.py --tag s100k --samples 100000 --vars 400000 --reps 3 >> gpurun_out/s40_streaming.jsonl 2>>gpurun_out/s40_streaming.err
UTMOS_B200_TRANSPOSE=1 python tools/bench_streaming.py --tag s100k_tr1 --samples 100000 --vars 400000 --reps 3 >> gpurun_out/s40_streaming.jsonl 2>>gpurun_out/s40_streaming.err
cat gpurun_out/s40_streaming.jsonl | cut -c1-400
python bench.py > gpurun_out/bench_s40.json 2> gpurun_out/bench_s40.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s40_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/s40_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"ingest_packed_kernel|transpose_bits" -c 2 -o gpurun_out/s40_streaming -f python tools/bench_streaming.py --reps 1 > gpurun_out/s40_ncu_full.log 2>&1; echo "ncu full rc=$?"
