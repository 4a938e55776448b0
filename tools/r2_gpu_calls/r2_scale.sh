# round 2: bench.py at N GPUs (argument), with the single-GPU verification leg
N=$1
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_scale_n$N.json 2> gpurun_out/r2_scale_n$N.err; echo "bench n$N rc=$?"; tail -c 300 gpurun_out/r2_scale_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_scale_n$N.json").read().strip().splitlines()[-1])
print(d["n_gpus"], "value", round(d["value"],2), "ms", round(d["ms_per_step"],3), "verified", d["verified_vs_single_gpu"], d["select_parts_ms"], "e2e ms", round(d["e2e"]["ms_per_step"],3), d["e2e"]["phases_ms"]["h2d_ms"])
PY
