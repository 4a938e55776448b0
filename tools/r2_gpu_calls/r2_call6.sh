# round 2, GPU call 6 (1 GPU): ncu full capture of K2 flavour 3, launch list of the bench
export UTMOS_B200_INGEST=3 UTMOS_B200_INGEST_TILE=55296
python tools/bench_streaming.py --reps 1 --tag plain > gpurun_out/r2_ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ingest_packed -c 1 -o gpurun_out/r2_ingest3 python tools/bench_streaming.py --reps 1 --tag ncu > gpurun_out/r2_ncu_ingest3.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2_ncu_ingest3.log
