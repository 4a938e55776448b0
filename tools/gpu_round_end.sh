# ingest look-back window A/B (1, 2, 4 tiles per lane) + the ingest stress test
: > gpurun_out/s42_streaming.jsonl
python tools/bench_streaming.py --tag look1 >> gpurun_out/s42_streaming.jsonl 2>gpurun_out/s42_streaming.err
UTMOS_B200_INGEST=2 python tools/bench_streaming.py --tag look2 >> gpurun_out/s42_streaming.jsonl 2>>gpurun_out/s42_streaming.err
python tools/bench_streaming.py --tag look1_again >> gpurun_out/s42_streaming.jsonl 2>>gpurun_out/s42_streaming.err
cut -c1-250 gpurun_out/s42_streaming.jsonl
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ingest_many or ragged or empty_and" > gpurun_out/s42_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/s42_pytest.log
