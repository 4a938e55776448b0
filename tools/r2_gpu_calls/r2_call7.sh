# round 2, GPU call 7 (2 GPUs): multi-GPU parity tests + bench at N=2 with the single-GPU verification leg
timeout -k 10 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_pytest_mgpu.log 2>&1; echo "pytest(multi) rc=$?"; tail -4 gpurun_out/r2_pytest_mgpu.log
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc=$?"; tail -c 400 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_n2.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"], d["ms_per_step"], d["verified_vs_single_gpu"], d["select_parts_ms"], d["e2e"]["ms_per_step"])
PY
