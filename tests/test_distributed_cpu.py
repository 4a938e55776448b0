"""world_size-2 gloo tests (CPU) of the host side of the variant-sharded path: shard split, handle all-gather,
gain / var_count all-reduce, and that a sharded run reproduces the single-process oracle.

The CUDA kernel cannot run here; a CPU stand-in with the DeviceMatrix surface plays one rank's shard and does the
per-step delta exchange through the same process group (the real kernel does it over NVLink)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import select_oracle as orc
from tests import helpers as H
from utmos_b200 import _native
from utmos_b200.distributed import HostCollectives, ShardedMatrix, shard_bounds


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 1000, 1103547):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (_, e), (b, _) in zip(spans[:-1], spans[1:]):
                assert e == b
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


class CpuShard:
    """DeviceMatrix stand-in: one rank's rows on the CPU, gains replicated, deltas summed over the group."""

    def __init__(self, n_samples, af_mode, rows_hint=0, device=0, flags=0):
        self.n, self.af_mode = n_samples, af_mode
        self.parts, self.afs = [], []
        self.exported = self.connected = self.layout = None

    def append_packed(self, gt, af=None):
        self.parts.append(np.asarray(gt))
        self.afs.append(None if af is None else np.asarray(af).reshape(-1))

    def rows(self):
        packed = np.concatenate(self.parts) if self.parts else np.zeros((0, (self.n + 7) // 8), np.uint8)
        self.keep, self.vc = orc.filter_rows_c(packed, self.n)
        self.packed = packed[self.keep]
        return int(self.keep.sum())

    def set_option(self, option, value):
        assert option == 4
        self.global_rows = value

    def finalize(self):
        self.dense = np.unpackbits(self.packed, axis=1, count=self.n).astype(np.int64)
        self.gain0 = self.dense.sum(axis=0).astype(np.uint32)
        return self.vc

    def info(self):
        return {"has_sample_major": 1}

    def mgpu_layout(self, row_base, merged_rows, allow_tail=True):
        assert row_base % 32 == 0 and merged_rows % 32 == 0 and merged_rows >= row_base and allow_tail
        self.layout = (row_base, merged_rows)

    def mgpu_export(self, rank, world):
        assert self.layout is not None                        # layout before export
        self.exported = (rank, world)
        return np.full(64, rank, dtype=np.uint8)

    def mgpu_connect(self, handles):
        assert handles.shape == (self.exported[1], 64)
        assert [int(h[0]) for h in handles] == list(range(self.exported[1]))
        self.connected = True

    def get_gains0(self):
        return self.gain0, np.zeros(self.n, np.uint64), np.zeros(self.n, np.uint64)

    def set_gains0(self, cnt, lo, hi, global_rows):
        self.gain0 = np.asarray(cnt, dtype=np.int64)
        self.global_rows = global_rows

    def begin(self, mask, weights=None):
        self.mask = np.array(mask, dtype=np.uint8)
        self.w = weights
        self.gain = self.gain0.astype(np.int64).copy()
        self.live = np.ones(self.dense.shape[0], dtype=bool)
        self.tot = 0

    def steps(self, max_steps):
        idx, new, score = [], [], []
        stop = 0
        for _ in range(max_steps):
            s = np.where(self.mask == 1, self.gain.astype(np.float64), 0.0)
            if self.w is not None:
                s = s * self.w
            best = int(np.argmax(s))
            if s[best] == 0:
                stop = 1
                break
            idx.append(best), new.append(int(self.gain[best])), score.append(float(s[best]))
            self.mask[best] = 0
            self.tot += int(self.gain[best])
            if self.tot >= self.global_rows:
                stop = 2
                break
            hit = self.live & (self.dense[:, best] == 1)
            delta = torch.from_numpy(self.dense[hit].sum(axis=0))
            dist.all_reduce(delta)                              # what the kernel does over NVLink
            self.gain -= delta.numpy()
            self.live &= ~hit
        return np.array(idx, np.int64), np.array(new, np.int64), np.array(score), stop

    def close(self):
        pass


def _worker(rank, world, port, out_queue):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    comm = HostCollectives()
    # collectives on numpy arrays
    summed = comm.all_reduce_sum(np.array([rank + 1, 10], dtype=np.uint32))
    wrap = comm.all_reduce_sum(np.array([2**63 + 5], dtype=np.uint64))
    gathered = comm.all_gather_bytes(np.full(64, rank, dtype=np.uint8))
    # sharded run on the reference fixtures
    parts = H.load_jl_parts(["chunk0.jl", "chunk1.jl"])
    n = 2504
    calls = []
    gather = comm.all_gather_bytes
    comm.all_gather_bytes = lambda payload: (calls.append(len(payload)), gather(payload))[1]
    sm = ShardedMatrix(n, _native.AF_NONE, comm=comm, matrix_factory=CpuShard)
    for part in parts:
        b, e = shard_bounds(part["GT"].shape[0], rank, world)
        sm.append_packed(part["GT"][b:e])
    var_count = sm.finalize()
    count_mode_gathers = list(calls)
    # --af: the fixed-point scale needs the global row count BEFORE the local finalize -> one more all-gather
    del calls[:]
    sa = ShardedMatrix(n, _native.AF_F64, comm=comm, matrix_factory=CpuShard)
    for part in parts:
        b, e = shard_bounds(part["GT"].shape[0], rank, world)
        sa.append_packed(part["GT"][b:e], part["AF"][b:e])
    vc_af = sa.finalize()
    af_mode_gathers = list(calls)
    af_rows_known_early = sa.local.global_rows == sm.num_vars and np.array_equal(vc_af, var_count)
    sa.close()
    comm.all_gather_bytes = gather
    sm.begin(np.ones(n, np.uint8))
    idx, new, score, stop = sm.steps(120)
    if rank == 0:
        out_queue.put((summed.tolist(), int(wrap[0]), gathered[:, 0].tolist(), sm.num_vars, var_count.tolist(),
                       idx.tolist(), new.tolist(), stop, sm.local.connected, count_mode_gathers, af_mode_gathers,
                       bool(af_rows_known_early)))
    sm.close()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_run_matches_oracle():
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, queue)) for r in range(world)]
    for p in procs:
        p.start()
    got = queue.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    summed, wrap, gathered, num_vars, var_count, idx, new, stop, connected, count_gathers, af_gathers, af_early = got
    n = 2504
    # count mode: ONE all-gather for gains + var_count + flags + row count, one for the IPC handles; --af: the row counts first
    assert count_gathers == [8 * (2 * n + 2), 64]
    assert af_gathers == [8, 8 * (4 * n + 2), 64] and af_early
    assert summed == [3, 20]
    assert wrap == (2 * (2**63 + 5)) % 2**64
    assert gathered == [0, 1] and connected
    gold = H.golden_json("full_order_count.json")
    names = np.asarray(H.load_jl_parts(["chunk0.jl"])[0]["samples"]).astype(str)
    assert num_vars == 1989
    assert [names[i] for i in idx] == [g[0] for g in gold["rows"][:120]]
    assert new == [g[2] for g in gold["rows"][:120]]
    assert [var_count[i] for i in idx] == [g[1] for g in gold["rows"][:120]]
    assert stop == 0
