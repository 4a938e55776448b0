"""Multi-GPU parity (needs >= 2 B200s on the node; skipped otherwise): rows sharded over 2 ranks must give the
same report rows and scores as the single-GPU path and as the recorded reference output."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import helpers as H
from utmos_b200 import _native, synth
from utmos_b200.distributed import HostCollectives, ShardedMatrix, shard_bounds

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, queue, use_af):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    comm = HostCollectives()
    out = {}
    # (1) reference fixtures, full ordering
    parts = H.load_jl_parts(["chunk0.jl", "chunk1.jl"])
    n = 2504
    sm = ShardedMatrix(n, _native.AF_F64 if use_af else _native.AF_NONE, device=rank, comm=comm)
    for part in parts:
        b, e = shard_bounds(part["GT"].shape[0], rank, world)
        sm.append_packed(part["GT"][b:e], part["AF"][b:e])
    vc = sm.finalize()
    sm.begin(np.ones(n, np.uint8))
    idx, new, score, stop = sm.steps(n)
    out["fixture"] = (sm.num_vars, vc.tolist(), idx.tolist(), new.tolist(), score.tolist(), stop, sm.info()["flavour"])
    sm.close()
    # (2) synthetic cohort with weights and exclusions, two batches of steps
    n_vars, n_samples = 40000, 1003
    gt, af = synth.mirror_rows(5, 0, n_vars, n_samples)
    wts = synth.synthetic_weights(n_samples)
    mask = np.ones(n_samples, np.uint8)
    mask[::97] = 2
    sm = ShardedMatrix(n_samples, _native.AF_F64 if use_af else _native.AF_NONE, device=rank, comm=comm)
    b, e = shard_bounds(n_vars, rank, world)
    sm.append_packed(gt[b:e], af[b:e])
    sm.finalize()
    sm.begin(mask, wts)
    i1, n1, s1, _ = sm.steps(40)
    i2, n2, s2, _ = sm.steps(60)
    out["synth"] = (np.concatenate([i1, i2]).tolist(), np.concatenate([n1, n2]).tolist(),
                    np.concatenate([s1, s2]).tolist())
    sm.close()
    # (3) a cohort beyond 65,535 samples: 32-bit carriers, sliced cluster tail on the merged lists
    n_vars, n_samples = 6000, 70001
    coh = synth.DeviceCohort(13, n_vars, n_samples, device=rank)
    gt, af = coh.to_host()
    coh.close()
    mask = np.ones(n_samples, np.uint8)
    mask[3::211] = 2
    sm = ShardedMatrix(n_samples, _native.AF_F64 if use_af else _native.AF_NONE, device=rank, comm=comm)
    b, e = shard_bounds(n_vars, rank, world)
    sm.append_packed(gt[b:e], af[b:e])
    sm.finalize()
    sm.begin(mask)
    i1, n1, s1, _ = sm.steps(300)
    out["wide"] = (i1.tolist(), n1.tolist(), s1.tolist(), sm.info()["flavour"])
    sm.close()
    if rank == 0:
        queue.put(out)
    dist.destroy_process_group()


@pytest.mark.parametrize("use_af", [False, True], ids=["count", "af"])
def test_two_gpus_match_single_gpu_and_reference(use_af):
    if _native.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, queue, use_af)) for r in range(world)]
    for p in procs:
        p.start()
    out = queue.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    gold = H.golden_json("full_order_af.json" if use_af else "full_order_count.json")
    names = np.asarray(H.load_jl_parts(["chunk0.jl"])[0]["samples"]).astype(str)
    num_vars, vc, idx, new, score, stop, flavour = out["fixture"]
    assert flavour == 5 and num_vars == 1989          # multi-GPU head + replicated tail
    assert [names[i] for i in idx] == [g[0] for g in gold["rows"]]
    assert new == [g[2] for g in gold["rows"]]
    assert [vc[i] for i in idx] == [g[1] for g in gold["rows"]]
    np.testing.assert_allclose(score, gold["argmax_scores"][:len(score)], rtol=1e-9 if use_af else 0)
    # single-GPU run of the synthetic case
    n_vars, n_samples = 40000, 1003
    gt, af = synth.mirror_rows(5, 0, n_vars, n_samples)
    wts = synth.synthetic_weights(n_samples)
    mask = np.ones(n_samples, np.uint8)
    mask[::97] = 2
    dm = _native.DeviceMatrix(n_samples, _native.AF_F64 if use_af else _native.AF_NONE)
    dm.append_packed(gt, af)
    dm.finalize()
    dm.begin(mask, wts)
    i1, n1, s1, _ = dm.steps(100)
    dm.close()
    assert out["synth"][0] == i1.tolist() and out["synth"][1] == n1.tolist() and out["synth"][2] == s1.tolist()
    # single-GPU run of the wide cohort
    n_vars, n_samples = 6000, 70001
    coh = synth.DeviceCohort(13, n_vars, n_samples)
    mask = np.ones(n_samples, np.uint8)
    mask[3::211] = 2
    dm = _native.DeviceMatrix(n_samples, _native.AF_F64 if use_af else _native.AF_NONE)
    dm.append_packed_device(coh.rows.ptr, n_vars, coh.pitch, coh.af.ptr)
    dm.finalize()
    dm.begin(mask)
    i1, n1, s1, _ = dm.steps(300)
    dm.close()
    coh.close()
    assert out["wide"][3] == 5
    assert out["wide"][0] == i1.tolist() and out["wide"][1] == n1.tolist() and out["wide"][2] == s1.tolist()


CLI_CASES = [
    ("select_af.txt", ["-c", "20", "--af", "chunk0.jl", "chunk1.jl"]),
    ("select_exclude.txt", ["-c", "20", "--exclude", "NA21117", "chunk0.jl", "chunk1.jl"]),
    ("select_first.txt", ["--maxmem", "1", "tiny.hdf5"]),
    ("select_af_h5.txt", ["--maxmem", "1", "-c", "20", "--lowmem", "tiny.af.hdf5"]),
]


@pytest.mark.parametrize("key,argv", CLI_CASES, ids=[c[0] + ":" + " ".join(c[1]) for c in CLI_CASES])
def test_cli_under_torchrun_reproduces_answer_keys(key, argv, tmp_path):
    """`torchrun --nproc-per-node 2 -m utmos_b200 select ...`: rows of every input sharded over two GPUs, rank 0
    writes the report; byte-identical to the reference's answer keys."""
    if _native.device_count() < 2:
        pytest.skip("needs two GPUs")
    import subprocess
    import sys
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "report.txt"
    args = [a if not a.endswith((".jl", ".hdf5")) else H.fixture(a) for a in argv]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), "-m", "utmos_b200", "select", "-o", str(out)] + args
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=240, check=False)
    assert res.returncode == 0, res.stderr[-2000:]
    assert out.read_text() == H.answer_key(key)
