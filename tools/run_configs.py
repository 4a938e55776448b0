#!/usr/bin/env python3
"""Runs the BASELINE.json configurations that are not bench.py's line (C3, C4, C5) through the public host API and
checks size-independent properties at the shapes the oracle cannot reach.  One JSON line per configuration.

  c3   2,504 x 1,103,547, --af + --weights + --subset + --exclude, --count -1        (1 GPU)
  c4   UKB-like S=50,000: a synthetic hdf5 file of the utmos dialect streamed chunk by chunk (--lowmem path)
  c5   gnomAD-like S=100,000: rows sharded over the ranks (torchrun) or one GPU

Properties checked (no oracle at these sizes):
  * winning scores never increase from one pick to the next (greedy on a submodular gain);
  * picks are unique, selectable (subset / exclude respected); sum(new_count) == tot_captured <= num_vars;
  * a second, independent kernel flavour (cluster kernel only, no list-driven tail) returns the same rows, bit for bit;
  * c4: the hdf5-streamed matrix returns the same rows as the same cohort ingested from packed .jl rows;
  * c5: written to --out so that the 1-GPU and N-GPU runs of the same cohort can be diffed (tools/diff_runs.py).
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from utmos_b200 import _native, synth  # noqa: E402  pylint: disable=wrong-import-position
from utmos_b200 import select as usel  # noqa: E402  pylint: disable=wrong-import-position


def check_run(idx, new, score, stop, mask, num_vars):
    assert len(set(idx.tolist())) == len(idx), "a sample was picked twice"
    assert np.all(mask[idx] == 1), "picked a sample that is not selectable"
    assert np.all(np.diff(score) <= 0), "winning scores increased"
    assert np.all(new >= 0) and int(new.sum()) <= num_vars
    if stop == _native.STOP_ALL:
        assert int(new.sum()) >= num_vars
    return {"steps": int(len(idx)), "stop": int(stop), "tot_captured": int(new.sum()), "num_vars": int(num_vars),
            "first_picks": idx[:5].tolist(), "first_scores": score[:5].tolist()}


def timed_selection(dm, mask, weights, count):
    _native.timer_start(0)
    dm.begin(mask, weights)
    idx, new, score, stop = dm.steps(count)
    return idx, new, score, stop, _native.timer_stop(0)


def config_c3(args):
    n_vars, n_samples = args.vars or 1_103_547, 2504
    cohort = synth.DeviceCohort(args.seed, n_vars, n_samples)
    names = synth.sample_names(n_samples)
    weights = synth.synthetic_weights(n_samples)
    subset = names[: n_samples // 2].tolist()
    exclude = names[::97].tolist()
    out = {"config": "c3", "samples": n_samples, "variants": n_vars, "flags": "--af --weights --subset --exclude --count -1"}
    runs = {}
    for label, flags in (("tail", 0), ("cluster_only", _native.F_NO_TAIL)):
        dm = _native.DeviceMatrix(n_samples, _native.AF_F64, rows_hint=n_vars, flags=flags)
        if args.tail_rows >= 0:
            dm.set_option(3, args.tail_rows)
        _native.timer_start(0)
        dm.append_packed_device(cohort.rows.ptr, n_vars, cohort.pitch, cohort.af.ptr)
        var_count = dm.finalize()
        load_ms = _native.timer_stop(0)
        data = usel.LoadedData(samples=names.astype("S"), data=dm, var_count=var_count)
        import pandas as pd
        wdf = pd.Series(weights, index=names, name="weight")
        _native.timer_start(0)
        t0 = time.perf_counter()
        rows = list(usel.run_selection(data, -1, subset, exclude, wdf))      # the public entry point (report rows)
        sel_ms = _native.timer_stop(0)
        wall = time.perf_counter() - t0
        # the raw columns again (scores are not part of the report)
        mask = np.where(np.isin(names, subset), 1, 2).astype(np.uint8)
        mask = np.where(np.isin(names, exclude), 2, mask).astype(np.uint8)
        idx, new, score, stop, _ = timed_selection(dm, mask, weights, n_samples)
        assert [r[0] for r in rows] == names[idx].tolist() and [r[2] for r in rows] == new.tolist()
        info = check_run(idx, new, score, stop, mask, dm.num_vars)
        info.update({"load_ms": load_ms, "select_ms": sel_ms, "timings": dm.timings(), "counters": dm.counters().tolist(), "host_wall_ms": wall * 1e3, "af_inexact": dm.info()["af_inexact"],
                     "report_tail": rows[-1][1:]})
        runs[label] = (idx, new, score, info)
        out[label] = info
        dm.close()
    a, b = runs["tail"], runs["cluster_only"]
    out["flavours_agree_bit_for_bit"] = bool(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]))
    assert out["flavours_agree_bit_for_bit"]
    return out


def config_c4(args):
    """hdf5 streaming.  Under torchrun every rank streams its own run of chunks of the same file (rank 0 writes it)."""
    n_samples = args.samples or 50_000
    n_vars = args.vars or 200_000
    from utmos_b200 import h5lite
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    comm = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
        from utmos_b200.distributed import HostCollectives
        comm = HostCollectives()
    names = synth.sample_names(n_samples).astype("S")
    path = os.path.join(args.tmp or tempfile.gettempdir(), f"utmos_c4_{n_samples}x{n_vars}.hdf5")
    mask = np.ones(n_samples, dtype=np.uint8)
    count = usel.resolve_select_count(args.count, n_samples)
    out = {"config": "c4", "samples": n_samples, "variants": n_vars, "dense_bytes": n_samples * n_vars, "count": args.count,
           "n_gpus": world}
    idx2 = new2 = stop2 = None
    if rank == 0:
        # the cohort: generated in HBM, mirrored to the host as packed .jl rows
        cohort = synth.DeviceCohort(args.seed, n_vars, n_samples, device=local_rank)
        gt, af = cohort.to_host()
        cohort.close()
        # (1) reference point: the same cohort ingested from packed rows on one GPU
        dm = _native.DeviceMatrix(n_samples, _native.AF_NONE, rows_hint=n_vars, device=local_rank)
        dm.append_packed(gt)
        var_count = dm.finalize()
        idx2, new2, score2, stop2, sel2_ms = timed_selection(dm, mask, None, count)
        out["packed"] = {"select_ms": sel2_ms, "timings": dm.timings(), "flavour": dm.info()["flavour"]}
        dm.close()
        # (2) write the hdf5 file of the utmos dialect (bool 'data', LZF chunks of max(1, int(1e6/4/S)) rows)
        t0 = time.perf_counter()
        writer = h5lite.H5Writer(path, names, float_data=False)
        block = max(1, (256 << 20) // n_samples)
        for r0 in range(0, n_vars, block):
            writer.append_packed(gt[r0:r0 + block], af[r0:r0 + block])
        writer.close(var_count)
        out["write_s"] = time.perf_counter() - t0
        out["hdf5_bytes"] = os.path.getsize(path)
        del gt, af
    if comm is not None:
        comm.barrier()
    # (3) the --lowmem path: stream the file chunk by chunk (LZF decode on the host, pinned staging, side stream)
    _native.timer_start(local_rank)
    t0 = time.perf_counter()
    data = usel.load_files([path], lowmem=1, device=local_rank, comm=comm)
    load_wall = time.perf_counter() - t0
    load_ms = _native.timer_stop(local_rank)
    dm = data["data"]
    local = dm.local if comm is not None else dm
    _native.timer_start(local_rank)
    dm.begin(mask, None)
    idx, new, score, stop = dm.steps(count)
    sel_ms = _native.timer_stop(local_rank)
    if rank == 0:
        assert np.array_equal(np.asarray(data["var_count"]), var_count)
        out["hdf5"] = check_run(idx, new, score, stop, mask, dm.shape[0])
        out["hdf5"].update({"load_wall_s": load_wall, "load_device_ms": load_ms, "select_ms": sel_ms, "timings": local.timings(),
                            "dense_GBps_streamed": n_samples * n_vars / 1e9 / load_wall, "flavour": local.info()["flavour"]})
        out["hdf5_and_packed_agree_bit_for_bit"] = bool(np.array_equal(idx, idx2) and np.array_equal(new, new2) and stop == stop2)
        assert out["hdf5_and_packed_agree_bit_for_bit"]
    data.close()
    if comm is not None:
        comm.barrier()
        import torch.distributed as dist
        dist.destroy_process_group()
    if rank == 0 and not args.keep:
        os.unlink(path)
    return out if rank == 0 else None


def config_c5(args):
    n_samples = args.samples or 100_000
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    total_vars = args.vars or 2_000_000
    from utmos_b200.distributed import shard_bounds
    begin, end = shard_bounds(total_vars, rank, world)
    comm = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        from utmos_b200.distributed import HostCollectives, ShardedMatrix
        comm = HostCollectives()
        sm = ShardedMatrix(n_samples, _native.AF_NONE, rows_hint=end - begin, device=local_rank, comm=comm)
        dm = sm.local
    else:
        sm = dm = _native.DeviceMatrix(n_samples, _native.AF_NONE, rows_hint=end - begin, device=local_rank)
    pitch = (n_samples + 7) // 8
    chunk_rows = max(1, (2 << 30) // pitch)             # generate 2 GiB of .jl-layout rows at a time in HBM
    tabs = synth.tables(n_samples)
    _native.timer_start(local_rank)
    t0 = time.perf_counter()
    for r0 in range(begin, end, chunk_rows):
        n = min(chunk_rows, end - r0)
        rows = _native.DeviceBuffer(n * pitch, local_rank)
        af = _native.DeviceBuffer(n * 8, local_rank)
        _native.synth_packed_device(args.seed, r0, n, n_samples, tabs[0], tabs[1], rows, af, local_rank)
        dm.append_packed_device(rows.ptr, n, pitch, None)
        dm.rows()                                         # synchronise before the generator buffers go away
        rows.close()
        af.close()
    var_count = sm.finalize()
    load_ms = _native.timer_stop(local_rank)
    load_wall = time.perf_counter() - t0
    mask = np.ones(n_samples, dtype=np.uint8)
    count = usel.resolve_select_count(args.count, n_samples)
    _native.timer_start(local_rank)
    sm.begin(mask)
    idx, new, score, stop = sm.steps(count)
    sel_ms = _native.timer_stop(local_rank)
    num_vars = sm.shape[0]
    info = check_run(idx, new, score, stop, mask, num_vars)
    out = {"config": "c5", "samples": n_samples, "variants": total_vars, "n_gpus": world, "rows_this_rank": end - begin,
           "count": args.count, "load_ms": load_ms, "load_wall_s": load_wall, "select_ms": sel_ms,
           "us_per_step": sel_ms * 1e3 / max(1, len(idx)), "packed_GB": total_vars * pitch / 1e9,
           "flavour": dm.info()["flavour"], "has_sample_major": dm.info()["has_sample_major"],
           "device_bytes": dm.info()["device_bytes"], "var_count_sum": int(np.asarray(var_count).sum())}
    out.update(info)
    if args.out and rank == 0:
        np.savez(args.out, idx=idx, new=new, score=score, stop=stop, var_count=np.asarray(var_count))
    sm.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return out if rank == 0 else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["c3", "c4", "c5"])
    ap.add_argument("--vars", type=int, default=0)
    ap.add_argument("--samples", type=int, default=0)
    ap.add_argument("--count", type=float, default=-1)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--tmp", default=None)
    ap.add_argument("--out", default=None)
    ap.add_argument("--keep", action="store_true")
    ap.add_argument("--tail-rows", type=int, default=-1, help="override the tail hand-over threshold (c3)")
    args = ap.parse_args()
    out = {"c3": config_c3, "c4": config_c4, "c5": config_c5}[args.config](args)
    if out is not None:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
