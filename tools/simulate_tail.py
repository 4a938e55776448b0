#!/usr/bin/env python3
"""CPU model of the greedy loop on the synthetic cohort (no GPU, NumPy only; ~4 minutes and ~1.5 GB at the full
1kGP shape).  It produced the statistics DESIGN.md quotes for the tail of a selection:

  * per regime of picks: rows newly covered per pick, share of them that only the picked sample carries
    (single-carrier rows need no edge-list entry), how often the exact runner-up's gain is untouched by the pick;
  * chained picks: with the exact top K per argmax round and picks taken down that list while the next candidate's
    gain has not moved, how many argmax rounds a full selection needs -- and that the pick order is exactly the
    plain greedy order (asserted), including the two-level (per-warp, then merged) top-K selection of
    select_tail_chain_kernel.

  python tools/simulate_tail.py [--vars 1103547] [--samples 2504] [--chain 1 2 4 8]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from utmos_b200 import synth  # noqa: E402  pylint: disable=wrong-import-position

REGIMES = ((0, 12), (12, 88), (88, 400), (400, 1000), (1000, 2000), (2000, 1 << 30))


def cohort(n_vars, n_samples, seed=0, step=100_000):
    parts = [synth.mirror_rows(seed, r0, min(step, n_vars - r0), n_samples)[0] for r0 in range(0, n_vars, step)]
    gt = np.concatenate(parts)
    cols = np.concatenate([np.packbits(np.unpackbits(gt[r0:r0 + step], axis=1, count=n_samples).T, axis=1)
                           for r0 in range(0, n_vars, step)], axis=1)
    gains = np.zeros(n_samples, np.int64)
    for r0 in range(0, n_vars, step):
        gains += np.unpackbits(gt[r0:r0 + step], axis=1, count=n_samples).sum(axis=0, dtype=np.int64)
    return gt, cols, gains


def two_level_top_k(g, k):
    """Element i belongs to thread i % 1024, warp (i % 1024) // 32: per-warp top-k lists, then a k-way merge."""
    n = len(g)
    warp = (np.arange(n) % 1024) // 32
    lists = []
    for w in range(32):
        m = np.flatnonzero(warp == w)
        o = m[np.lexsort((m, -g[m]))][:k]
        lists.append([(int(g[i]), int(i)) for i in o] + [(0, 0x7fffffff)] * (k - len(o)))
    ptr, top = [0] * 32, []
    for _ in range(k):
        heads = [lists[l][ptr[l]] if ptr[l] < k else (0, 0x7fffffff) for l in range(32)]
        mk = max(h[0] for h in heads)
        mi = min(h[1] for h in heads if h[0] == mk)
        for l in range(32):
            if heads[l][1] == mi and mi != 0x7fffffff:
                ptr[l] += 1
        top.append((mk, mi))
    return top


def run(gt, cols, gains0, n_samples, k):
    gains, live, used = gains0.copy(), np.full(cols.shape[1], 0xff, np.uint8), np.zeros(n_samples, bool)
    picks, rounds, stats = [], [], []
    while len(picks) < n_samples:
        g = np.where(used, 0, gains)
        top = two_level_top_k(g, max(k, 2))
        exact = np.lexsort((np.arange(n_samples), -g))[:max(k, 2)]
        assert [t[1] for t in top if t[0] > 0] == [int(i) for i in exact if g[i] > 0]
        if top[0][0] == 0:
            break
        taken = 0
        for c, (key, a) in enumerate(top[:k]):
            if c > 0 and (key == 0 or gains[a] != key):
                break
            new_mask = live & cols[a]
            live &= ~new_mask
            idx = np.flatnonzero(np.unpackbits(new_mask))
            sub = np.unpackbits(gt[idx], axis=1, count=n_samples)
            dec = sub.sum(axis=0, dtype=np.int64)
            if c == 0:
                runner = top[1][1]
                stats.append((len(idx), int((sub.sum(axis=1) == 1).sum()), int(runner != 0x7fffffff and dec[runner] == 0)))
            gains -= dec
            used[a] = True
            picks.append((a, key))
            taken += 1
        rounds.append((len(picks) - taken, taken))
    return picks, np.array(rounds), np.array(stats)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vars", type=int, default=1_103_547)
    ap.add_argument("--samples", type=int, default=2504)
    ap.add_argument("--chain", type=int, nargs="*", default=[1, 2, 4, 8])
    args = ap.parse_args()
    gt, cols, gains0 = cohort(args.vars, args.samples)
    ref = None
    for k in args.chain:
        picks, rounds, stats = run(gt, cols, gains0, args.samples, k)
        if ref is None:
            ref = picks
        assert picks == ref, "chained picks differ from plain greedy"
        line = [f"K={k}: {len(rounds)} argmax rounds for {len(picks)} picks;"]
        for lo, hi in REGIMES:
            m = (rounds[:, 0] >= lo) & (rounds[:, 0] < hi)
            if m.any():
                line.append(f"{lo}-{min(hi, len(picks))}: {rounds[m, 1].sum() / m.sum():.2f} picks/round")
        print("  ".join(line))
        if k == 1:
            for lo, hi in REGIMES:
                s = stats[lo:hi]
                if len(s):
                    print(f"  picks {lo}-{min(hi, len(stats))}: {s[:, 0].mean():.0f} rows per pick, single-carrier share "
                          f"{s[:, 1].sum() / max(1, s[:, 0].sum()):.2f}, runner-up untouched {s[:, 2].mean():.2f}")


if __name__ == "__main__":
    main()
