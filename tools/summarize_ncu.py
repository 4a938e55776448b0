#!/usr/bin/env python3
"""Text summaries of the ncu artefacts kept under profiles/ (run here, no GPU needed).

  summarize_ncu.py launches <launches.csv>         per-kernel totals and shares of a `--metrics gpu__time_duration.sum` pass
  summarize_ncu.py report <file.ncu-rep> [n]       per-launch duration / DRAM bytes / issue rate / stall mix, and the n
                                                   source lines with the most stall samples (needs -lineinfo builds)
"""
import collections
import csv
import io
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "").replace("utmos::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        agg[name][0] += 1
        agg[name][1] += v
    total = sum(v[1] for v in agg.values())
    print(f"{len(data)} launches, {total / 1e3:.3f} ms of kernel time (cold-cache, serialised: compare shares)")
    for name, (cnt, us) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{name:50s} {cnt:5d} launches {us:12.1f} us {100 * us / total:6.1f}%")


def ncu_csv(rep, *args):
    out = subprocess.run(["ncu", "-i", rep, "--csv", *args], capture_output=True, text=True, check=False).stdout
    return list(csv.reader(io.StringIO(out)))


def report(rep, top=25):
    rows = ncu_csv(rep, "--page", "raw")
    hdr, units = rows[0], rows[1]
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
            "smsp__inst_executed.sum", "sm__inst_executed.sum.per_cycle_elapsed", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "launch__cluster_size",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active"]
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")
             and "not_issued" not in h]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")][:110])
        for k in keys:
            if k in hdr:
                print(f"   {k:62s} {r[hdr.index(k)]} {units[hdr.index(k)]}")
        mix = sorted(((float(r[hdr.index(h)] or 0), h.split("stalled_")[1].replace("_per_issue_active.ratio", "")) for h in stall),
                     reverse=True)[:6]
        print("   stall mix (warps per issue): " + ", ".join(f"{n} {v:.2f}" for v, n in mix))
    src = ncu_csv(rep, "--page", "source", "--print-source", "cuda,sass")
    cur, hdr2, agg = None, None, collections.OrderedDict()
    for r in src:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr2 = r
        elif hdr2 and r[0] != "" and len(r) >= 8 and r[0].isdigit():
            try:
                key = (cur, int(r[0]))
                ent = agg.setdefault(key, [r[1], 0, 0])
                ent[1] += int(r[6])
                ent[2] += int(r[7])
            except ValueError:
                pass
    ts = sum(v[1] for v in agg.values()) or 1
    ti = sum(v[2] for v in agg.values()) or 1
    print(f"-- source lines by stall samples ({ts} samples, {ti} warp instructions, all captured launches)")
    for (f, line), (text, smp, inst) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print(f"   {f}:{line:<5d} samples {100 * smp / ts:5.1f}%  instr {100 * inst / ti:5.1f}%  {text.strip()[:90]}")


if __name__ == "__main__":
    if len(sys.argv) < 3:
        sys.exit(__doc__)
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        report(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
