run() { tag=$1; shift; env "$@" timeout -k 10 150 python bench.py --no-cpu --no-verify --steps 3 $EXTRA > gpurun_out/r2_b4_$tag.json 2> gpurun_out/r2_b4_$tag.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_b4_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["select_parts_ms"].items()}, "gain_ms", round(d["phases_ms"]["gain_ms"],3), d["phase_cycles"][5:8], d["phase_cycles"][12], "golden", d["verified_vs_oracle_golden"])
except Exception as e:
    print("$tag", "failed", e, open("gpurun_out/r2_b4_$tag.err").read()[-300:])
PY
}
EXTRA="--config c3"
run c3_old UTMOS_B200_LAZY=0
run c3_lazy0 UTMOS_B200_LAZY_G1=256 UTMOS_B200_LAZY_G4=2048
run c3_lazy1024 UTMOS_B200_LAZY_ROWS=1024 UTMOS_B200_LAZY_G1=256 UTMOS_B200_LAZY_G4=2048
run c3_lazy256 UTMOS_B200_LAZY_ROWS=256 UTMOS_B200_LAZY_G1=256 UTMOS_B200_LAZY_G4=2048
run c3_lazy256_s3 UTMOS_B200_LAZY_ROWS=256 UTMOS_B200_LAZY_G1=256 UTMOS_B200_LAZY_G4=2048 UTMOS_B200_LAZY_SLACK=3
EXTRA=""
run c2_old UTMOS_B200_LAZY=0
run c2_lazy256 UTMOS_B200_LAZY_ROWS=256 UTMOS_B200_LAZY_G1=256 UTMOS_B200_LAZY_G4=2048
run c2_lazy128_s8 UTMOS_B200_LAZY_ROWS=128 UTMOS_B200_LAZY_G1=256 UTMOS_B200_LAZY_G4=2048 UTMOS_B200_LAZY_SLACK=8
EXTRA="--tail-rows 4096"; run c2_old_tr4096 UTMOS_B200_LAZY=0
EXTRA="--tail-rows 1024"; run c2_old_tr1024 UTMOS_B200_LAZY=0
EXTRA="--single-rows 512"; run c2_old_sr512 UTMOS_B200_LAZY=0
EXTRA="--single-rows 1024 --tail-rows 4096"; run c2_old_sr1024_tr4096 UTMOS_B200_LAZY=0
python tools/bench_convert.py > gpurun_out/r2_convert_fast.json 2>&1; cat gpurun_out/r2_convert_fast.json
python tools/bench_convert.py --samples 2500 > gpurun_out/r2_convert_fast_tile.json 2>&1; cat gpurun_out/r2_convert_fast_tile.json
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "convert or read_vcf or af or c3" > gpurun_out/r2_pytest_c.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_c.log
