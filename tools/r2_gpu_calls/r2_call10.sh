# round 2, GPU call 10 (4 GPUs): owner-computes cluster flavour of the replicated tail, sweep of its hand-over threshold
run() { tag=$1; shift; timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 3 --warmup 2 --no-verify "$@" > gpurun_out/r2_b10_$tag.json 2> gpurun_out/r2_b10_$tag.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_b10_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(d["value"],2), round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["select_parts_ms"].items()}, "e2e", round(d["e2e"]["ms_per_step"],3))
except Exception as e:
    print("$tag", "failed", e, open("gpurun_out/r2_b10_$tag.err").read()[-400:])
PY
}
run default
run sr1024 --single-rows 1024
run sr4096 --single-rows 4096
run sr512 --single-rows 512
