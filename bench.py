#!/usr/bin/env python3
"""bench.py -- `utmos select --count -1` on the 1kGP chr22 shape (2,504 samples x 1,103,547 variants), count mode.

One "step" = one complete greedy selection over one synthetic cohort: packed .jl rows -> ingest (filter,
re-pitch, bit order) -> sample-major copy -> column reduce -> up to S greedy steps -> report columns.

  value   whole-job packed GB/s with the raw .jl-layout rows already resident in HBM
          (packed bytes V*ceil(S/8) summed over ranks / wall time of a step, max over ranks)
  e2e     same metric through the public host API (utmos_b200._native.DeviceMatrix over the C ABI) with the
          rows in pinned HOST memory: H2D copies and the D2H read of the report columns are in the timed region
  roofline  the dominant kernel (the persistent selection kernel): algorithmic bytes / CUDA-event duration
  cpu_baseline  the NumPy port of the reference's row loop (oracle/select_oracle.py DenseOracle), one thread,
          on a bounded sample, extrapolated to the full run and labelled so

`--impl reference` times that CPU port alone (the reference is pure Python + NumPy, single threaded).
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
WORKLOAD = "1kGP chr22 shape: 2,504 samples x 1,103,547 variants, count-based, --count -1"
METRIC = "utmos_select_packed_GBps"
UNIT = "GB/s"
CPU_SAMPLE_ROWS = 65536


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


def traffic_from_ncu():
    """DRAM bytes (read + write) the dominant kernel moved per selection, from the committed ncu capture
    (profiles/r1_traffic.json; cannot be measured live).  Only meaningful for the default single-GPU workload."""
    path = os.path.join(ROOT, "profiles", "r1_traffic.json")
    try:
        with open(path) as fh:
            return json.load(fh)["per_selection_bytes"]
    except (OSError, KeyError, ValueError):
        return None


class ClockSampler:
    """nvidia-smi clock / throttle-reason log during the timed region (B200_PROFILING.md)."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu_index), "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        try:
            with open(self.path) as fh:
                for line in fh:
                    f = [x.strip() for x in line.split(",")]
                    if len(f) < 9:
                        continue
                    try:
                        sm.append(float(f[1]))
                        smax.append(float(f[2]))
                    except ValueError:
                        continue
                    for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"],
                                         f[5:9]):
                        if val.lower().startswith("active"):
                            reasons.add(name)
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def select_bytes(n_vars, pitch, n_samples, steps, new_total):
    """Algorithmic bytes of the selection kernel: each newly covered row is read once
    (new_total * pitch), every step probes one column and updates the live mask (3*V/8) and scans the gains."""
    return new_total * pitch + steps * (3 * n_vars // 8 + 13 * n_samples)


def cpu_port_step_seconds(rows_gt, n_samples, n_steps):
    """Seconds per greedy step of the NumPy port on a dense bool slab (one thread)."""
    from oracle import select_oracle as orc
    dense = np.unpackbits(rows_gt, axis=1, count=n_samples).astype(bool)
    oracle = orc.DenseOracle(dense, np.ones(n_samples, dtype=np.uint8))
    t0 = time.perf_counter()
    for _ in range(n_steps):
        oracle.mask[oracle.score_once()[0]] = 0
    return (time.perf_counter() - t0) / n_steps


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (NumPy port) on a bounded sample, rank 0 only."""
    if rank != 0:
        return
    from utmos_b200 import synth
    pitch = (N_SAMPLES + 7) // 8
    gt, _af = synth.mirror_rows(args.seed, 0, CPU_SAMPLE_ROWS, N_SAMPLES)
    total = args.warmup + args.steps
    per_step = []
    from oracle import select_oracle as orc
    dense = np.unpackbits(gt, axis=1, count=N_SAMPLES).astype(bool)
    oracle = orc.DenseOracle(dense, np.ones(N_SAMPLES, dtype=np.uint8))
    for i in range(total):
        t0 = time.perf_counter()
        oracle.mask[oracle.score_once()[0]] = 0
        if i >= args.warmup:
            per_step.append(time.perf_counter() - t0)
    sec = float(np.mean(per_step))
    full_step = sec * N_VARS / CPU_SAMPLE_ROWS              # one greedy step over the whole matrix
    full_run = full_step * N_SAMPLES                         # --count -1: up to S greedy steps
    value = N_VARS * pitch / 1e9 / full_run
    sample = (f"{args.steps} greedy steps of the NumPy port of utmos/select.py:24-53 on the first {CPU_SAMPLE_ROWS} "
              f"rows (dense bool) of the same cohort; per-step time scaled by V/{CPU_SAMPLE_ROWS} and by S={N_SAMPLES} "
              "steps (extrapolated full run)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": full_run * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64-popcount/f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "measured_ms_per_greedy_step_on_sample": sec * 1e3},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--vars", type=int, default=N_VARS, help="rows per rank (default: the 1kGP chr22 count)")
    ap.add_argument("--flags", type=int, default=0, help="utmos_create flags (kernel flavour)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--step-times", action="store_true", help="record per-pick timestamps (profiling; slows the loop)")
    ap.add_argument("--tail-rows", type=int, default=-1, help="override the tail hand-over threshold")
    ap.add_argument("--regain-rows", type=int, default=-1, help="override the heavy-pick threshold (UTMOS_OPT_REGAIN_ROWS)")
    ap.add_argument("--single-rows", type=int, default=-1, help="override the cluster-tail -> single-CTA-tail threshold")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
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
    n_vars, n_samples = args.vars, N_SAMPLES
    pitch = (n_samples + 7) // 8
    packed_bytes = n_vars * pitch

    # synthetic cohort generated in HBM (each rank its own rows of the same seeded cohort), mirrored to pinned host
    cohort = synth.DeviceCohort(args.seed, n_vars, n_samples, device=device, row0=rank * n_vars)
    pinned = _native.PinnedBuffer(packed_bytes)
    cohort.rows.to_host(pinned.array)
    host_rows = pinned.array.reshape(n_vars, pitch)
    mask = np.ones(n_samples, dtype=np.uint8)

    comm = None
    if world > 1:
        from utmos_b200.distributed import HostCollectives, ShardedMatrix
        comm = HostCollectives()

    def one_selection(resident):
        t = [time.perf_counter()]
        if world > 1:
            sm = ShardedMatrix(n_samples, _native.AF_NONE, rows_hint=n_vars, device=device, flags=args.flags, comm=comm)
            dm = sm.local
        else:
            sm = dm = _native.DeviceMatrix(n_samples, _native.AF_NONE, rows_hint=n_vars, device=device, flags=args.flags)
        t.append(time.perf_counter())
        if resident:
            dm.append_packed_device(cohort.rows.ptr, n_vars, pitch, 0)
        else:
            dm.append_packed(host_rows, None)
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
        sm.begin(mask)
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
        outs = [one_selection(resident) for _ in range(args.steps)]
        wall = time.perf_counter() - t0              # every selection ends synchronised (results copied to host)
        elapsed = _native.timer_stop(device) / 1e3   # device synchronised again; event-to-event seconds
        wall_vs_event.append((wall, elapsed))
        barrier()
        if dist is not None:
            import torch
            t = torch.tensor([elapsed], device=f"cuda:{local_rank}", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            elapsed = float(t.item())
        return elapsed, outs

    sampler = ClockSampler(device)
    if rank == 0:
        sampler.start()
    t_res, outs_res = timed(True)
    t_e2e, outs_e2e = timed(False)
    clocks = sampler.stop() if rank == 0 else {}

    idx, new, score, stop, var_count, info, _ = outs_res[-1]
    idx2, new2 = outs_e2e[-1][0], outs_e2e[-1][1]
    assert np.array_equal(idx, idx2) and np.array_equal(new, new2), "resident and host paths disagree"

    ms_res = t_res / args.steps * 1e3
    ms_e2e = t_e2e / args.steps * 1e3
    value = world * packed_bytes / 1e9 / (t_res / args.steps)
    e2e_value = world * packed_bytes / 1e9 / (t_e2e / args.steps)

    # device-side (CUDA event) time of each phase, averaged over the timed resident steps
    phases = {k: float(np.mean([o[6][k] for o in outs_res])) for k in outs_res[0][6] if k not in ("step_ns", "counters", "host_ms")}
    phases_e2e = {k: float(np.mean([o[6][k] for o in outs_e2e])) for k in outs_e2e[0][6] if k not in ("step_ns", "counters", "host_ms")}
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
                           "select_tail_kernel (head: select_cluster_kernel + regain_kernel)", "select_mgpu_kernel",
                           "select_tail_kernel replicated on every rank (head: select_mgpu_kernel, hand-over: build_edges_kernel over NVLink)"][info["flavour"]],
                "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic_from_ncu() if world == 1 and n_vars == N_VARS else None, "algorithmic_bytes": sel_bytes, "kernel_ms": phases["select_ms"],
                "note": "latency-bound: %d dependent greedy steps, %.2f us per step; achieved = algorithmic bytes of the whole greedy loop "
                        "(SURVEY.md 8d: newly covered rows once + per step a column probe, live-mask update and gain scan) / "
                        "CUDA-event time of all its launches; traffic = ncu DRAM bytes of the tail kernel's launches of one "
                        "selection (the lists replace the per-step column probe, so traffic is far below the algorithmic bytes)" %
                        (n_steps, phases["select_ms"] * 1e3 / max(n_steps, 1))}
    # one-time streaming kernels against the same peak
    row_bytes = info["num_vars"] // world * info["row_pitch_bytes"]
    streaming = {}
    if phases["ingest_ms"] > 0:
        streaming["ingest"] = {"bytes": packed_bytes + row_bytes, "ms": phases["ingest_ms"],
                               "frac": (packed_bytes + row_bytes) / 1e9 / (phases["ingest_ms"] / 1e3) / peak}
    if phases["transpose_ms"] > 0:
        streaming["transpose"] = {"bytes": 2 * row_bytes, "ms": phases["transpose_ms"],
                                  "frac": 2 * row_bytes / 1e9 / (phases["transpose_ms"] / 1e3) / peak}
    if phases["gain_ms"] > 0:
        streaming["colreduce"] = {"bytes": row_bytes, "ms": phases["gain_ms"],
                                  "frac": row_bytes / 1e9 / (phases["gain_ms"] / 1e3) / peak}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    cpu_baseline = None
    if not args.no_cpu and world == 1:
        from oracle import select_oracle as orc
        gt, _ = synth.mirror_rows(args.seed, 0, CPU_SAMPLE_ROWS, n_samples)
        t0 = time.perf_counter()
        sec = cpu_port_step_seconds(gt, n_samples, 12)
        full_run = sec * n_vars / CPU_SAMPLE_ROWS * n_steps
        cpu_baseline = {"value": packed_bytes / 1e9 / full_run, "unit": UNIT, "cores": 1, "kind": "port",
                        "sample": f"12 greedy steps of the NumPy port of utmos/select.py:24-53 (DenseOracle) on the first "
                                  f"{CPU_SAMPLE_ROWS} rows of the same cohort; scaled by V/{CPU_SAMPLE_ROWS} and by the "
                                  f"{n_steps} steps the GPU run took (extrapolated full run {full_run:.0f} s)",
                        "host_cores_available": os.cpu_count()}
        # the plain-C restatement, full --count -1 run on a reduced row count (measured, not extrapolated)
        c_rows = 16384
        t1 = time.perf_counter()
        c_idx, _, _, _ = orc.greedy_c(gt[:c_rows], n_samples, mask, None, None, n_samples)
        c_sec = time.perf_counter() - t1
        cpu_baseline["c_oracle"] = {"rows": c_rows, "steps": int(len(c_idx)), "seconds": c_sec,
                                    "note": "oracle/greedy_oracle.c full ordering on the first %d rows, 1 thread" % c_rows}
        cpu_baseline["seconds_spent"] = time.perf_counter() - t0

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_res, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32-popcount/i64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "samples": n_samples, "variants_per_gpu": n_vars, "seed": args.seed,
                       "parallelism": ("variants sharded over %d GPUs (%d rows each), gains replicated; head: per-step P2P delta exchange; tail: edge lists merged on every rank over NVLink, replicated single-CTA kernel" % (world, n_vars)) if world > 1 else "single GPU",
                       "l2": "inputs (345 MB packed + 353 MB sample-major copy) exceed the 126 MB L2",
                       "greedy_steps": n_steps, "stop": int(stop), "flags": args.flags},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(packed_bytes + n_samples),
                    "d2h_bytes_per_step": int(3 * 8 * n_steps + 8 * n_samples + 64), "phases_ms": phases_e2e},
            "gpu_launches": int(info["kernel_launches"]) * args.steps,
            "roofline": roofline, "streaming_kernels": streaming, "phases_ms": phases,
            "cpu_baseline": cpu_baseline, "clocks": clocks,
            "step_profile": step_profile, "host_ms_create_append_finalize_begin_steps_close": {"resident": np.mean([o[6]["host_ms"] for o in outs_res], axis=0).round(3).tolist(), "e2e": np.mean([o[6]["host_ms"] for o in outs_e2e], axis=0).round(3).tolist()}, "phase_cycles": [int(x) for x in outs_res[-1][6]["counters"][:13]], "timing": {"clock": "CUDA events around the K timed steps (utmos_timer_start/stop: device synchronised on both sides), max over ranks", "host_wall_s_resident_e2e": [round(w, 6) for w, _ in wall_vs_event], "event_s_resident_e2e": [round(e, 6) for _, e in wall_vs_event]}, "select_time_s": phases["select_ms"] / 1e3, "us_per_greedy_step": phases["select_ms"] * 1e3 / max(n_steps, 1)}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
