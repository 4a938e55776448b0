"""Variant-sharded selection over the GPUs of one node: one process per GPU.

Rows (variants) are split over the ranks; every rank keeps the full, replicated per-sample gain vectors.  The
host side only does plumbing with ``torch.distributed`` (gloo for the small host tensors): it all-gathers the
CUDA IPC handles of the exchange blocks, all-reduces the step-0 gains / var_count / row counts, and then every
rank calls ``steps`` with the same arguments.  The per-step exchange itself happens inside the CUDA kernel
(``select_mgpu_kernel``: P2P stores of the gain deltas into the peers' inboxes over NVLink).
"""
import numpy as np

from utmos_b200 import _native


def shard_bounds(n_rows, rank, world):
    """Contiguous, balanced split of ``n_rows`` rows: [begin, end) of ``rank``."""
    base, extra = divmod(int(n_rows), int(world))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


class HostCollectives:
    """Small host-side collectives on numpy arrays over a gloo process group."""

    def __init__(self, group=None):
        import torch
        import torch.distributed as dist
        self._torch, self._dist = torch, dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        if group is None and dist.get_backend() != "gloo":
            group = dist.new_group(backend="gloo")          # host tensors need a CPU-capable backend
        self.group = group
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()

    def all_reduce_sum(self, arr):
        """Element-wise sum over ranks; unsigned 64-bit arrays wrap modulo 2^64 like the device limbs do."""
        arr = np.ascontiguousarray(arr)
        if arr.dtype == np.uint64:
            t = self._torch.from_numpy(arr.view(np.int64).copy())
            self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM, group=self.group)
            return t.numpy().view(np.uint64)
        t = self._torch.from_numpy(arr.astype(np.int64))
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM, group=self.group)
        return t.numpy().astype(arr.dtype if arr.dtype != np.uint32 else np.int64)

    def all_gather_bytes(self, payload):
        """payload: uint8[n] -> uint8[world, n] in rank order."""
        mine = self._torch.from_numpy(np.ascontiguousarray(payload, dtype=np.uint8).copy())
        out = [self._torch.empty_like(mine) for _ in range(self.world)]
        self._dist.all_gather(out, mine, group=self.group)
        return np.stack([o.numpy() for o in out])

    def barrier(self):
        self._dist.barrier(group=self.group)


class ShardedMatrix:
    """Same surface as ``_native.DeviceMatrix`` for one rank's shard of the rows.

    ``shape`` / ``var_count`` / ``steps`` report GLOBAL quantities (identical on every rank), so
    ``utmos_b200.select.run_selection`` can drive it unchanged.
    """

    def __init__(self, n_samples, af_mode=_native.AF_NONE, rows_hint=0, device=0, flags=0, comm=None,
                 matrix_factory=None):
        self.comm = comm if comm is not None else HostCollectives()
        factory = matrix_factory or _native.DeviceMatrix
        self.local = factory(n_samples, af_mode, rows_hint=rows_hint, device=device, flags=flags)
        self.n_samples = int(n_samples)
        self.af_mode = af_mode
        self.num_vars = None
        self.var_count = None

    # ingestion of THIS rank's rows
    def append_packed(self, gt, af=None):
        self.local.append_packed(gt, af)

    def append_packed2(self, gt2, af=None):
        self.local.append_packed2(gt2, af)

    def append_packed_device(self, d_rows, n_rows, pitch, d_af=None):
        self.local.append_packed_device(d_rows, n_rows, pitch, d_af)

    def append_dense(self, chunk):
        self.local.append_dense(chunk)

    def append_h5_chunks(self, *args, **kwargs):
        self.local.append_h5_chunks(*args, **kwargs)


    def finalize(self):
        comm = self.comm
        kept = self.local.rows()
        with_af = self.af_mode != _native.AF_NONE
        rows_all = None
        if with_af or comm.world == 1:
            # AF limbs are fixed point: every rank must use the scale of the WHOLE matrix before it reduces its columns
            rows_all = comm.all_gather_bytes(np.array([kept], dtype=np.int64).view(np.uint8)).view(np.int64).reshape(-1)
            self.local.set_option(4, int(rows_all.sum()))
        local_vc = self.local.finalize()
        local_vc = np.asarray(local_vc, dtype=np.int64)
        if comm.world > 1:
            # ONE all-gather carries everything the ranks have to sum: step-0 gains (counts, and the AF limbs when there
            # are any), var_count, "I have a sample-major copy" and the rank's row count; every rank adds the shares up in
            # rank order, so the sums are identical everywhere.  (Six gloo collectives per finalize used to cost ~5 ms on
            # 8 ranks; count mode needs nothing from its peers before this point.)
            cnt, lo, hi = self.local.get_gains0()
            has_cols = int(self.local.info()["has_sample_major"])
            parts = [cnt.astype(np.int64), local_vc, np.array([has_cols, kept], dtype=np.int64)]
            if with_af:
                parts += [lo.view(np.int64), hi.view(np.int64)]
            mine = np.concatenate(parts)
            shares = comm.all_gather_bytes(mine.view(np.uint8)).view(np.int64).reshape(comm.world, -1)
            with np.errstate(over="ignore"):
                summed = shares.sum(axis=0, dtype=np.int64)              # limbs wrap modulo 2^64 like on the device
            n = self.n_samples
            cnt_sum, vc_sum, cols_sum = summed[:n], summed[n:2 * n], int(summed[2 * n])
            rows_all = shares[:, 2 * n + 1].copy()
            total = int(rows_all.sum())
            if with_af:
                lo = summed[2 * n + 2:3 * n + 2].view(np.uint64)
                hi = summed[3 * n + 2:4 * n + 2].view(np.uint64)
            self.local.set_gains0(cnt_sum.astype(np.uint32), lo, hi, total)
            # merged row numbering for the hand-over to the replicated tail (every rank's rows padded to 32)
            padded = (rows_all + 31) // 32 * 32
            self.local.mgpu_layout(int(padded[:comm.rank].sum()), int(padded.sum()), cols_sum == comm.world)
            handle = self.local.mgpu_export(comm.rank, comm.world)
            self.local.mgpu_connect(comm.all_gather_bytes(handle))
            self.var_count = vc_sum.copy()
        else:
            total = int(rows_all.sum())
            self.var_count = comm.all_reduce_sum(local_vc)
        self.num_vars = total
        comm.barrier()
        return self.var_count

    @property
    def shape(self):
        return (self.num_vars, self.n_samples)

    @property
    def dtype(self):
        return self.local.dtype

    def begin(self, mask, weights=None):
        self.local.begin(mask, weights)
        self.comm.barrier()                                   # every rank's state is reset before any kernel starts

    def steps(self, max_steps):
        return self.local.steps(max_steps)

    def info(self):
        return self.local.info()

    def timings(self, reset=False):
        return self.local.timings(reset)

    def close(self):
        self.comm.barrier()                                   # no rank unmaps its inbox while a peer may still write
        self.local.close()
