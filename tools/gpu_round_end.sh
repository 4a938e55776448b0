#!/bin/bash
# First GPU call of the next round (about 11 GPU-minutes; run it with --timeout 900): the opt-in heavy-pick kernel through the whole suite, then
# the sweep that decides its defaults.  Everything lands in gpurun_out/.
#   1. full parity suite with cover_decrement_kernel forced on (count mode; AF flavours keep regain_kernel)
#   2. bench lines: default, decrement, decrement with lower thresholds (the mid picks, 1,800-4,500 rows each, then take
#      the kernel too) and more queued head launches per host check
UTMOS_B200_DECREMENT=1 python -m pytest tests -m gpu -x -q > gpurun_out/next_pytest_decrement.log 2>&1; echo "pytest(decrement) rc=$?"; tail -2 gpurun_out/next_pytest_decrement.log
#   3. chained picks in the tail (select_tail_chain_kernel, exact top-K per argmax round): parity subset + bench lines
for k in 2 4; do
  UTMOS_B200_TAIL_CHAIN=$k python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_orderings or random_cases or synthetic_reduced or step_batches or tail_flavours or cli_answer" > gpurun_out/next_pytest_chain$k.log 2>&1; echo "pytest(chain $k) rc=$?"; tail -2 gpurun_out/next_pytest_chain$k.log
  UTMOS_B200_TAIL_CHAIN=$k python bench.py --no-cpu > gpurun_out/next_bench_chain$k.json 2> gpurun_out/next_bench_chain$k.err
  UTMOS_B200_TAIL_CHAIN=$k UTMOS_B200_DECREMENT=1 python bench.py --no-cpu > gpurun_out/next_bench_chain${k}_dec.json 2> gpurun_out/next_bench_chain${k}_dec.err
done
#   4. edge lists without single-carrier rows
UTMOS_B200_SKIP_SINGLE=1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_orderings or random_cases or synthetic_reduced or step_batches or tail_flavours or cli_answer or float32_af or cohorts_whose" > gpurun_out/next_pytest_skipsingle.log 2>&1; echo "pytest(skip single) rc=$?"; tail -2 gpurun_out/next_pytest_skipsingle.log
UTMOS_B200_SKIP_SINGLE=1 python bench.py --no-cpu > gpurun_out/next_bench_skipsingle.json 2> gpurun_out/next_bench_skipsingle.err
UTMOS_B200_SKIP_SINGLE=1 UTMOS_B200_TAIL_CHAIN=4 UTMOS_B200_DECREMENT=1 python bench.py --no-cpu > gpurun_out/next_bench_all_optins.json 2> gpurun_out/next_bench_all_optins.err
#   5. K2 ingest with batched loads (flavour 2): parity subset + streaming micro-bench A/B
UTMOS_B200_INGEST=2 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ingest_many or ragged or empty_and or synthetic_reduced or cli_answer" > gpurun_out/next_pytest_ingest2.log 2>&1; echo "pytest(ingest 2) rc=$?"; tail -2 gpurun_out/next_pytest_ingest2.log
: > gpurun_out/next_streaming.jsonl
python tools/bench_streaming.py --tag ingest1 >> gpurun_out/next_streaming.jsonl
UTMOS_B200_INGEST=2 python tools/bench_streaming.py --tag ingest2 >> gpurun_out/next_streaming.jsonl
UTMOS_B200_INGEST=2 UTMOS_B200_INGEST_TILE=55296 python tools/bench_streaming.py --tag ingest2_tile54k >> gpurun_out/next_streaming.jsonl
UTMOS_B200_INGEST=2 python tools/bench_streaming.py --tag ingest2_s100k --samples 100000 --vars 400000 --reps 3 >> gpurun_out/next_streaming.jsonl
cut -c1-260 gpurun_out/next_streaming.jsonl
#   6. hdf5 chunks decoded on the GPU (lzf_unpack_bool_kernel): the hdf5 parity tests, then config C4 both ways
UTMOS_B200_H5_GPU_LZF=1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hdf5 or lowmem or h5 or cli_answer" > gpurun_out/next_pytest_gpulzf.log 2>&1; echo "pytest(gpu lzf) rc=$?"; tail -2 gpurun_out/next_pytest_gpulzf.log
python tools/run_configs.py c4 --vars 200000 > gpurun_out/next_c4_host_lzf.json 2> gpurun_out/next_c4_host_lzf.err
UTMOS_B200_H5_GPU_LZF=1 python tools/run_configs.py c4 --vars 200000 > gpurun_out/next_c4_gpu_lzf.json 2> gpurun_out/next_c4_gpu_lzf.err
tail -c 600 gpurun_out/next_c4_host_lzf.json; echo; tail -c 600 gpurun_out/next_c4_gpu_lzf.json; echo
python bench.py --no-cpu > gpurun_out/next_bench_default.json 2> gpurun_out/next_bench_default.err
for cfg in "4096 4" "2048 8" "1024 8" "512 16" "256 16"; do
  set -- $cfg
  UTMOS_B200_DECREMENT=1 UTMOS_B200_HEAD_REPS=$2 python bench.py --no-cpu --regain-rows $1 > gpurun_out/next_bench_dec_$1_$2.json 2> gpurun_out/next_bench_dec_$1_$2.err
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/next_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 3), round(d["phases_ms"]["select_ms"], 3), d["gpu_launches"])
    except Exception as e:  # noqa
        print(f, "failed", e)
PY
