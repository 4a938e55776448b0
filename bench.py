#!/usr/bin/env python3
"""bench.py -- `utmos select --count -1` on the 1kGP chr22 shape (2,504 samples x 1,103,547 variants), count mode.

One "step" = one complete greedy selection over one synthetic cohort: packed .jl rows -> ingest (filter,
re-pitch, bit order) -> sample-major copy -> column reduce -> up to S greedy steps -> report columns.

  value   whole-job packed GB/s with the raw .jl-layout rows already resident in HBM
          (packed bytes V*ceil(S/8) summed over ranks / wall time of a step, max over ranks)
  e2e     same metric through the public host API (utmos_b200._native.DeviceMatrix over the C ABI) with the
          rows in HOST memory: H2D copies and the D2H read of the report columns are in the timed region
  roofline  the dominant kernel family (the greedy loop): algorithmic bytes / CUDA-event duration
  cpu_baseline  the UNMODIFIED reference (oracle/_ref, installed from /root/reference by build()) timed to
          completion on a reduced row count on this box's host cores, next to this repo's e2e path on the same rows
  verified_vs_oracle_golden  the picks of the last timed selection equal tests/golden/c2_full_order.npz (plain-C oracle
          run at the full shape in the build container)
  verified_vs_single_gpu  (N > 1) rank 0 re-runs the whole N-shard cohort alone and gets the same rows
  configs.c3  the same shape with --af --weights --subset --exclude (fixed-point float64 path), oracle-checked

`--impl reference` times the reference's own CPU implementation alone (single threaded: that is all it uses).
`--config c3` makes C3 the timed workload of the line.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SAMPLES = 2504
N_VARS = 1_103_547
WORKLOADS = {
    "c2": "1kGP chr22 shape: 2,504 samples x 1,103,547 variants, count-based, --count -1",
    "c3": "1kGP chr22 shape: 2,504 samples x 1,103,547 variants, --af --weights --subset --exclude, --count -1",
}
METRIC = "utmos_select_packed_GBps"
UNIT = "GB/s"
CPU_REF_ROWS = 2048          # rows of the complete reference run (about 9 s on one core)
CPU_C_ROWS = 16384


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


def traffic_from_ncu():
    """DRAM bytes (read + write) the dominant kernel moved per selection, from the committed ncu capture
    (profiles/r1_traffic.json; cannot be measured live).  Only meaningful for the default single-GPU workload."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as fh:
                return json.load(fh)["per_selection_bytes"]
        except (OSError, KeyError, ValueError):
            continue
    return None


class ClockSampler:
    """nvidia-smi clock / throttle-reason log during the timed region (B200_PROFILING.md)."""

    QUERY = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu_index), "-lms", "20"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    @staticmethod
    def _stamp(text):
        import datetime
        try:
            return datetime.datetime.strptime(text.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

    def stop(self, windows=()):
        """windows: (t0, t1) wall-clock pairs of the timed regions; samples inside them are the ones reported (the sampler
        itself runs from the start of the program, nvidia-smi needs a few hundred ms to deliver its first line)."""
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = []
        try:
            with open(self.path) as fh:
                for line in fh:
                    f = [x.strip() for x in line.split(",")]
                    if len(f) < 9:
                        continue
                    try:
                        rows.append((self._stamp(f[0]), float(f[1]), float(f[2]),
                                     [name for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                                                 "sw_power_cap"], f[5:9]) if val.lower().startswith("active")]))
                    except ValueError:
                        continue
            os.unlink(self.path)
        except OSError:
            pass
        inside = [r for r in rows if r[0] is not None and any(a - 0.02 <= r[0] <= b + 0.02 for a, b in windows)]
        use = inside if len(inside) >= 3 else rows
        if use:
            out = {"sm_mhz": float(np.median([r[1] for r in use])), "sm_max_mhz": float(max(r[2] for r in use)),
                   "reasons": sorted({name for r in use for name in r[3]}), "samples": len(use),
                   "samples_inside_timed_regions": len(inside), "samples_whole_run": len(rows), "period_ms": 20}
        return out


def select_bytes(n_vars, pitch, n_samples, steps, new_total):
    """Algorithmic bytes of the selection loop: each newly covered row is read once
    (new_total * pitch), every step probes one column and updates the live mask (3*V/8) and scans the gains."""
    return new_total * pitch + steps * (3 * n_vars // 8 + 13 * n_samples)


def c3_options(n_samples):
    """--weights / --subset / --exclude of config C3: mask (1 selectable, 2 excluded) and weights."""
    from utmos_b200 import synth
    names = synth.sample_names(n_samples)
    weights = synth.synthetic_weights(n_samples)
    mask = np.where(np.isin(names, names[: n_samples // 2]), 1, 2).astype(np.uint8)
    mask = np.where(np.isin(names, names[::97]), 2, mask).astype(np.uint8)
    return mask, weights


def golden(config, seed, n_vars, n_samples):
    """The plain-C oracle's full ordering for this exact cohort (tests/golden/<config>_full_order.npz) or None."""
    path = os.path.join(ROOT, "tests", "golden", f"{config}_full_order.npz")
    if not os.path.exists(path):
        return None
    g = np.load(path)
    if int(g["seed"]) != seed or int(g["n_vars"]) != n_vars or int(g["n_samples"]) != n_samples:
        return None
    return g


# ------------------------------------------------------------------------------------------------------------
# the reference's CPU implementation (oracle/_ref = the unmodified package; else the NumPy port)
# ------------------------------------------------------------------------------------------------------------
def reference_runner():
    """(callable(dense bool matrix, names) -> number of report rows, kind)."""
    import logging
    try:
        from oracle import refloader
        ref = refloader.load_reference_select()
        logging.disable(logging.CRITICAL)           # the reference logs one INFO line per pick

        def run(dense, names):
            data = {"data": dense, "samples": names.astype("S"), "var_count": dense.sum(axis=0)}
            return len(list(ref.run_selection(data, -1, None, None, None)))
        return run, "reference"
    except Exception:  # pylint: disable=broad-except
        from oracle import select_oracle as orc

        def run(dense, names):
            return len(orc.DenseOracle(dense, np.ones(dense.shape[1], dtype=np.uint8)).run(dense.shape[1])[0])
        return run, "port"


def reference_sample(seed, rows):
    from utmos_b200 import synth
    gt, _af = synth.mirror_rows(seed, 0, rows, N_SAMPLES)
    dense = np.unpackbits(gt, axis=1, count=N_SAMPLES).astype(bool)
    return gt, dense, synth.sample_names(N_SAMPLES)


REF_NOTE = ("complete --count -1 selection (utmos/select.py run_selection -> greedy_select -> calculate_scores) on the first "
            "%d rows of the same seeded cohort, dense bool matrix as load_files builds it, one thread (the reference is "
            "single threaded).  Measured, not extrapolated.  The reference's packed GB/s FALLS as rows are added (more "
            "greedy steps, each a pass over the uncovered rows: 9.1e-5 / 7.0e-5 / 4.3e-5 / 1.9e-5 GB/s at 1,024 / 2,048 / "
            "4,096 / 20,000 rows in the build container), so this value is an upper bound of its throughput at the full "
            "1,103,547 rows (where one run takes about 4.5 h) and a speed-up computed from it is a lower bound")


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation on a bounded sample, rank 0 only."""
    if rank != 0:
        return
    pitch = (N_SAMPLES + 7) // 8
    run, kind = reference_runner()
    _gt, dense, names = reference_sample(args.seed, args.ref_rows)
    per_step, picks = [], 0
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        picks = run(dense, names)
        if i >= args.warmup:
            per_step.append(time.perf_counter() - t0)
    sec = float(np.mean(per_step))
    value = args.ref_rows * pitch / 1e9 / sec
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64/i64 (NumPy)", "data": "synthetic",
            "config": {"workload": WORKLOADS["c2"]}, "run_config": {"sample_rows": args.ref_rows, "picks": picks},
            "extrapolated": False,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": REF_NOTE % args.ref_rows,
                             "host_cores_available": os.cpu_count()},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3"], help="timed workload (BASELINE.json configs[1] / [2])")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--vars", type=int, default=N_VARS, help="rows per rank (default: the 1kGP chr22 count)")
    ap.add_argument("--ref-rows", type=int, default=CPU_REF_ROWS, help="rows of the complete reference run (CPU baseline)")
    ap.add_argument("--flags", type=int, default=0, help="utmos_create flags (kernel flavour)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-verify", action="store_true", help="skip the N>1 single-GPU verification leg and the C3 leg")
    ap.add_argument("--pageable", action="store_true", help="e2e from pageable host rows (what joblib.load hands the CLI) instead of pinned")
    ap.add_argument("--step-times", action="store_true", help="record per-pick timestamps (profiling; slows the loop)")
    ap.add_argument("--tail-rows", type=int, default=-1, help="override the tail hand-over threshold")
    ap.add_argument("--regain-rows", type=int, default=-1, help="override the heavy-pick threshold (UTMOS_OPT_REGAIN_ROWS)")
    ap.add_argument("--single-rows", type=int, default=-1, help="override the cluster-tail -> single-CTA-tail threshold")
    ap.add_argument("--list-budget", type=int, default=-1, help="override the edge-list budget (entries) of the list-driven tail")
    ap.add_argument("--heavy-rows", type=int, default=-1, help="override the entry-divided-cluster -> shared-memory tail threshold")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    from utmos_b200 import _native, synth
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    device = local_rank
    sampler = ClockSampler(device)
    if rank == 0:
        sampler.start()                              # early: nvidia-smi needs a few hundred ms before its first sample
    timed_windows = []
    n_vars, n_samples = args.vars, N_SAMPLES
    pitch = (n_samples + 7) // 8
    packed_bytes = n_vars * pitch
    with_af = args.config == "c3"
    af_mode = _native.AF_F64 if with_af else _native.AF_NONE

    # synthetic cohort generated in HBM (each rank its own rows of the same seeded cohort), mirrored to host memory
    cohort = synth.DeviceCohort(args.seed, n_vars, n_samples, device=device, row0=rank * n_vars)
    pinned = _native.PinnedBuffer(packed_bytes)
    cohort.rows.to_host(pinned.array)
    host_rows = pinned.array.reshape(n_vars, pitch)
    if args.pageable:
        host_rows = np.array(host_rows, copy=True)
    host_af = None
    if with_af:
        host_af = np.empty(n_vars, dtype=np.float64)
        cohort.af.to_host(host_af.view(np.uint8))
    if with_af:
        mask, weights = c3_options(n_samples)
    else:
        mask, weights = np.ones(n_samples, dtype=np.uint8), None

    comm = None
    if world > 1:
        from utmos_b200.distributed import HostCollectives, ShardedMatrix
        comm = HostCollectives()

    def one_selection(resident, rows_dev=None, rows_host=None, rows_n=None, af_dev=None, single=False, af_mode_=None,
                      mask_=None, weights_=None, flags_=None):
        """One complete selection.  Default inputs: this rank's cohort; `single` forces the one-GPU code path."""
        rows_n = n_vars if rows_n is None else rows_n
        mode = af_mode if af_mode_ is None else af_mode_
        msk = mask if mask_ is None else mask_
        wts = weights if weights_ is None and mask_ is None else weights_
        t = [time.perf_counter()]
        fl = args.flags if flags_ is None else flags_
        if world > 1 and not single:
            sm = ShardedMatrix(n_samples, mode, rows_hint=rows_n, device=device, flags=fl, comm=comm)
            dm = sm.local
        else:
            sm = dm = _native.DeviceMatrix(n_samples, mode, rows_hint=rows_n, device=device, flags=fl)
        if args.list_budget >= 0:
            dm.set_option(11, args.list_budget)
        t.append(time.perf_counter())
        if resident:
            parts = rows_dev if rows_dev is not None else [(cohort.rows.ptr, n_vars, cohort.af.ptr)]
            for ptr, n, afp in parts:
                dm.append_packed_device(ptr, n, pitch, afp if mode != _native.AF_NONE else 0)
        else:
            dm.append_packed(host_rows if rows_host is None else rows_host, host_af if mode != _native.AF_NONE else None)
        t.append(time.perf_counter())
        var_count = sm.finalize()
        t.append(time.perf_counter())
        if args.step_times:
            dm.set_option(2, 1)
        if args.tail_rows >= 0:
            dm.set_option(3, args.tail_rows)
        if args.single_rows >= 0:
            dm.set_option(5, args.single_rows)
        if args.regain_rows >= 0:
            dm.set_option(1, args.regain_rows)
        if args.heavy_rows >= 0:
            dm.set_option(10, args.heavy_rows)
        sm.begin(msk, wts)
        t.append(time.perf_counter())
        idx, new, score, stop = sm.steps(n_samples)
        t.append(time.perf_counter())
        info, tim = dm.info(), dm.timings()
        info["num_vars"] = sm.shape[0]
        tim["step_ns"] = dm.step_times(0, len(idx))
        tim["counters"] = dm.counters()
        sm.close()
        t.append(time.perf_counter())
        tim["host_ms"] = [round((b - a) * 1e3, 3) for a, b in zip(t[:-1], t[1:])]
        return idx, new, score, stop, var_count, info, tim

    wall_vs_event = []

    def timed(resident):
        for _ in range(args.warmup):
            one_selection(resident)
        barrier()
        _native.timer_start(device)                  # device synchronised, CUDA event recorded
        t0 = time.perf_counter()
        w0 = time.time()
        outs = [one_selection(resident) for _ in range(args.steps)]
        wall = time.perf_counter() - t0              # every selection ends synchronised (results copied to host)
        elapsed = _native.timer_stop(device) / 1e3   # device synchronised again; event-to-event seconds
        timed_windows.append((w0, time.time()))
        wall_vs_event.append((wall, elapsed))
        barrier()
        if dist is not None:
            import torch
            t = torch.tensor([elapsed], device=f"cuda:{local_rank}", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            elapsed = float(t.item())
        return elapsed, outs

    t_res, outs_res = timed(True)
    t_e2e, outs_e2e = timed(False)
    clocks = sampler.stop(timed_windows) if rank == 0 else {}

    idx, new, score, stop, var_count, info, _ = outs_res[-1]
    idx2, new2 = outs_e2e[-1][0], outs_e2e[-1][1]
    assert np.array_equal(idx, idx2) and np.array_equal(new, new2), "resident and host paths disagree"

    # ---- parity legs (untimed) ---------------------------------------------------------------------------
    verified_golden = None
    if world == 1:
        gold = golden(args.config, args.seed, n_vars, n_samples)
        if gold is not None:
            ok = (np.array_equal(idx, gold["idx"]) and np.array_equal(new, gold["new"]) and stop == int(gold["stop"])
                  and np.array_equal(var_count, gold["var_count"]) and np.array_equal(score, gold["score"]))
            assert ok, "the timed selection differs from the plain-C oracle's full ordering (tests/golden)"
            verified_golden = True
    verified_single = None
    if world > 1 and not args.no_verify:
        # rank 0 alone, one-GPU code path, the whole N-shard cohort: must give the same rows (utmos/select.py:48 first-index
        # argmax, :97-112) as the sharded run.  The other ranks wait at the barrier.
        if rank == 0:
            shards = [synth.DeviceCohort(args.seed, n_vars, n_samples, device=device, row0=r * n_vars) for r in range(world)]
            ref_out = one_selection(True, rows_dev=[(s.rows.ptr, n_vars, s.af.ptr) for s in shards], rows_n=n_vars * world,
                                    single=True)
            for s in shards:
                s.close()
            same = (np.array_equal(ref_out[0], idx) and np.array_equal(ref_out[1], new) and ref_out[3] == stop
                    and np.array_equal(ref_out[4], var_count) and np.array_equal(ref_out[2], score))
            assert same, "N-GPU selection differs from the single-GPU selection of the same cohort"
            verified_single = True
        barrier()

    ms_res = t_res / args.steps * 1e3
    ms_e2e = t_e2e / args.steps * 1e3
    value = world * packed_bytes / 1e9 / (t_res / args.steps)
    e2e_value = world * packed_bytes / 1e9 / (t_e2e / args.steps)

    # device-side (CUDA event) time of each phase, averaged over the timed resident steps
    skip = ("step_ns", "counters", "host_ms")
    phases = {k: float(np.mean([o[6][k] for o in outs_res])) for k in outs_res[0][6] if k not in skip}
    phases_e2e = {k: float(np.mean([o[6][k] for o in outs_e2e])) for k in outs_e2e[0][6] if k not in skip}
    step_ns = outs_res[-1][6]["step_ns"].astype(np.float64)
    gaps = np.diff(step_ns) / 1e3                      # us between consecutive picks
    step_profile = {}
    if args.step_times and len(gaps) > 200:
        step_profile = {"first_20_steps_ms": float(gaps[:20].sum() / 1e3), "steps_20_200_ms": float(gaps[20:200].sum() / 1e3),
                        "steps_200_end_ms": float(gaps[200:].sum() / 1e3),
                        "tail_us_per_step_median": float(np.median(gaps[200:])),
                        "tail_us_per_step_p90": float(np.percentile(gaps[200:], 90)),
                        "max_us": float(gaps.max())}
    if args.step_times and rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "step_series.json"), "w") as fh:
            json.dump({"gap_us": gaps.tolist(), "new": new.tolist()}, fh)
    peak, peak_kind = peaks()
    n_steps = len(idx)
    sel_bytes = select_bytes(info["num_vars"], info["row_pitch_bytes"], n_samples, n_steps, int(new.sum())) // world
    achieved = sel_bytes / 1e9 / (phases["select_ms"] / 1e3) if phases["select_ms"] > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": ["argmax_step_kernel+cover_step_kernel", "select_persistent_kernel", "select_cluster_kernel",
                           "select_tail_kernel (head: select_cluster_kernel + cover_decrement_kernel)", "select_mgpu_kernel",
                           "select_tail_kernel replicated on every rank (head: select_mgpu_kernel, hand-over: build_edges_kernel over NVLink)"][info["flavour"]],
                "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic_from_ncu() if world == 1 and n_vars == N_VARS and not with_af else None, "algorithmic_bytes": sel_bytes, "kernel_ms": phases["select_ms"],
                "note": "latency-bound: %d dependent greedy steps, %.2f us per step; achieved = algorithmic bytes of the whole greedy loop "
                        "(SURVEY.md 8d: newly covered rows once + per step a column probe, live-mask update and gain scan) / "
                        "CUDA-event time of all its launches; traffic = ncu DRAM bytes of the tail kernel's launches of one "
                        "selection (the lists replace the per-step column probe, so traffic is far below the algorithmic bytes)" %
                        (n_steps, phases["select_ms"] * 1e3 / max(n_steps, 1))}
    # one-time streaming kernels against the same peak
    row_bytes = info["num_vars"] // world * info["row_pitch_bytes"]
    af_bytes = 8 * n_vars if with_af else 0
    streaming = {}
    if phases["ingest_ms"] > 0:
        streaming["ingest"] = {"bytes": packed_bytes + row_bytes + 2 * af_bytes, "ms": phases["ingest_ms"],
                               "frac": (packed_bytes + row_bytes + 2 * af_bytes) / 1e9 / (phases["ingest_ms"] / 1e3) / peak}
    if phases["transpose_ms"] > 0:
        streaming["transpose"] = {"bytes": 2 * row_bytes, "ms": phases["transpose_ms"],
                                  "frac": 2 * row_bytes / 1e9 / (phases["transpose_ms"] / 1e3) / peak}
    if phases["gain_ms"] > 0:
        streaming["colreduce"] = {"bytes": row_bytes + af_bytes, "ms": phases["gain_ms"],
                                  "frac": (row_bytes + af_bytes) / 1e9 / (phases["gain_ms"] / 1e3) / peak}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- config C3 beside the line (single GPU, default workload) ----------------------------------------
    configs = {}
    if world == 1 and not with_af and not args.no_verify and n_vars == N_VARS:
        c3_mask, c3_w = c3_options(n_samples)
        c3_outs = []
        for i in range(2 + 3):
            if i == 2:
                _native.timer_start(device)
            c3_outs.append(one_selection(True, af_mode_=_native.AF_F64, mask_=c3_mask, weights_=c3_w))
        c3_ms = _native.timer_stop(device) / 3
        c_idx, c_new, c_score, c_stop, c_vc, c_info, _ = c3_outs[-1]
        c3_ph = {k: float(np.mean([o[6][k] for o in c3_outs[2:]])) for k in c3_outs[0][6] if k not in skip}
        gold3 = golden("c3", args.seed, n_vars, n_samples)
        c3_ok = None
        if gold3 is not None:
            c3_ok = bool(np.array_equal(c_idx, gold3["idx"]) and np.array_equal(c_new, gold3["new"]) and
                         np.array_equal(c_score, gold3["score"]) and c_stop == int(gold3["stop"]))
            assert c3_ok, "config C3 differs from the plain-C oracle's full ordering (tests/golden/c3_full_order.npz)"
        gain_bytes = c_info["num_vars"] * c_info["row_pitch_bytes"] + 8 * n_vars
        configs["c3"] = {"workload": WORKLOADS["c3"], "ms_per_selection": c3_ms, "picks": int(len(c_idx)), "stop": int(c_stop),
                         "select_ms": c3_ph["select_ms"], "us_per_greedy_step": c3_ph["select_ms"] * 1e3 / max(1, len(c_idx)),
                         "phases_ms": c3_ph, "af_inexact": c_info["af_inexact"], "verified_vs_oracle_golden": c3_ok,
                         "step0_gains": {"bytes": gain_bytes, "ms": c3_ph["gain_ms"],
                                         "frac": gain_bytes / 1e9 / (c3_ph["gain_ms"] / 1e3) / peak if c3_ph["gain_ms"] > 0 else None},
                         "dtype": "u64 fixed-point limbs -> f64 (one rounding per comparison)"}
        # the same selection in the REFERENCE's tie order (UTMOS_F_REF_TIES, what `utmos select --af` runs on one GPU):
        # checked against the plain-C oracle's mode-0 run (sequential float64 sums, utmos/select.py:37-41)
        gold3r = golden("c3ref", args.seed, n_vars, n_samples)
        if gold3r is not None:
            one_selection(True, af_mode_=_native.AF_F64, mask_=c3_mask, weights_=c3_w, flags_=_native.F_REF_TIES)
            _native.timer_start(device)
            r_idx, r_new, r_score, r_stop, _vc, _info, r_tim = one_selection(True, af_mode_=_native.AF_F64, mask_=c3_mask,
                                                                            weights_=c3_w, flags_=_native.F_REF_TIES)
            r_ms = _native.timer_stop(device)
            r_ok = bool(np.array_equal(r_idx, gold3r["idx"]) and np.array_equal(r_new, gold3r["new"]) and
                        r_stop == int(gold3r["stop"]) and np.allclose(r_score, gold3r["score"], rtol=1e-12, atol=0))
            assert r_ok, "config C3 with UTMOS_F_REF_TIES differs from the reference-order oracle (tests/golden/c3ref_full_order.npz)"
            configs["c3"]["reference_tie_order"] = {
                "ms_per_selection": r_ms, "select_ms": r_tim["select_ms"], "verified_vs_reference_order_golden": r_ok,
                "select_parts_ms": {"head": r_tim.get("head_ms"), "hand_over": r_tim.get("handover_ms"), "tail": r_tim.get("tail_ms")},
                "picks_in_the_same_position_as_the_exact_order": int(np.sum(r_idx == c_idx)) if len(r_idx) == len(c_idx) else None,
                "note": "UTMOS_F_REF_TIES: per-step kernels (argmax with the replay, cover, streaming recompute after heavy picks) until the edge lists can be built, then the entry-divided cluster tail with the replay of near-tie candidates inside (rows from the edge lists, sorted in shared memory, summed sequentially in float64)"}

    # ---- CPU baseline: the unmodified reference, complete run on a reduced row count, and this repo on the same rows
    cpu_baseline = None
    if not args.no_cpu and world == 1:
        t0 = time.perf_counter()
        run, kind = reference_runner()
        gt_small, dense, names = reference_sample(args.seed, args.ref_rows)
        t1 = time.perf_counter()
        ref_picks = run(dense, names)
        ref_sec = time.perf_counter() - t1
        # the same rows through this repo's host-buffer path (H2D inside), count mode, --count -1
        small_ms = []
        for i in range(6):
            if i >= 3:
                _native.timer_start(device)
            dm = _native.DeviceMatrix(n_samples, _native.AF_NONE, rows_hint=args.ref_rows, device=device)
            dm.append_packed(gt_small, None)
            dm.finalize()
            dm.begin(np.ones(n_samples, dtype=np.uint8))
            s_idx, _s_new, _s_score, _s_stop = dm.steps(n_samples)
            dm.close()
            if i >= 3:
                small_ms.append(_native.timer_stop(device))
        assert len(s_idx) == ref_picks, "reference and GPU path emit a different number of report rows on the reduced cohort"
        ref_value = args.ref_rows * pitch / 1e9 / ref_sec
        cpu_baseline = {"value": ref_value, "unit": UNIT, "cores": 1, "kind": kind, "sample": REF_NOTE % args.ref_rows,
                        "extrapolated": False, "host_cores_available": os.cpu_count(),
                        "like_for_like": {"rows": args.ref_rows, "picks": ref_picks, "reference_seconds": ref_sec,
                                          "this_repo_e2e_ms": float(np.mean(small_ms)),
                                          "measured_speedup": ref_sec * 1e3 / float(np.mean(small_ms))}}
        # the plain-C restatement, full --count -1 run on a reduced row count (measured)
        from oracle import select_oracle as orc
        gt_c, _ = synth.mirror_rows(args.seed, 0, CPU_C_ROWS, n_samples)
        t1 = time.perf_counter()
        c_idx2, _, _, _ = orc.greedy_c(gt_c, n_samples, np.ones(n_samples, dtype=np.uint8), None, None, n_samples)
        c_sec = time.perf_counter() - t1
        cpu_baseline["c_oracle"] = {"rows": CPU_C_ROWS, "steps": int(len(c_idx2)), "seconds": c_sec,
                                    "note": "oracle/greedy_oracle.c full ordering on the first %d rows, 1 thread; the full 1,103,547-row "
                                            "ordering took it 183 s in the build container (tests/golden/c2_full_order.npz)" % CPU_C_ROWS}
        cpu_baseline["seconds_spent"] = time.perf_counter() - t0

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_res, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64 fixed-point/f64" if with_af else "u32-popcount/i64", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.config]},
            "run_config": {"samples": n_samples, "variants_per_gpu": n_vars, "seed": args.seed,
                       "parallelism": ("variants sharded over %d GPUs (%d rows each), gains replicated; head: per-step P2P delta exchange; tail: edge lists merged on every rank over NVLink, replicated tail kernel" % (world, n_vars)) if world > 1 else "single GPU",
                       "l2": "inputs (345 MB packed + 353 MB sample-major copy) exceed the 126 MB L2",
                       "greedy_steps": n_steps, "stop": int(stop), "flags": args.flags,
                       "host_rows": "pageable" if args.pageable else "pinned"},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(packed_bytes + n_samples + (8 * n_vars + 8 * n_samples if with_af else 0)),
                    "d2h_bytes_per_step": int(3 * 8 * n_steps + 8 * n_samples + 64), "phases_ms": phases_e2e},
            "gpu_launches": int(info["kernel_launches"]) * args.steps,
            "verified_vs_oracle_golden": verified_golden, "verified_vs_single_gpu": verified_single,
            "roofline": roofline, "streaming_kernels": streaming, "phases_ms": phases,
            "select_parts_ms": {"head": phases.get("head_ms"), "hand_over": phases.get("handover_ms"), "tail": phases.get("tail_ms")},
            "configs": configs, "cpu_baseline": cpu_baseline, "clocks": clocks,
            "step_profile": step_profile, "host_ms_create_append_finalize_begin_steps_close": {"resident": np.mean([o[6]["host_ms"] for o in outs_res], axis=0).round(3).tolist(), "e2e": np.mean([o[6]["host_ms"] for o in outs_e2e], axis=0).round(3).tolist()}, "phase_cycles": [int(x) for x in outs_res[-1][6]["counters"][:16]], "timing": {"clock": "CUDA events around the K timed steps (utmos_timer_start/stop: device synchronised on both sides), max over ranks", "host_wall_s_resident_e2e": [round(w, 6) for w, _ in wall_vs_event], "event_s_resident_e2e": [round(e, 6) for _, e in wall_vs_event]}, "select_time_s": phases["select_ms"] / 1e3, "us_per_greedy_step": phases["select_ms"] * 1e3 / max(n_steps, 1)}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
