# opt-in cover_decrement kernel: parity subset + bench line with it switched on
UTMOS_B200_DECREMENT=1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "regain_threshold or synthetic_reduced or full_orderings" > gpurun_out/s44_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s44_pytest.log
UTMOS_B200_DECREMENT=1 python bench.py --no-cpu > gpurun_out/bench_s44_dec.json 2> gpurun_out/bench_s44_dec.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench_s44_dec.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['phases_ms'],d['gpu_launches'])"
