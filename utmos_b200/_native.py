"""ctypes binding of libutmos_b200.so (include/utmos_b200.h).  The only bridge to the CUDA code.

There is deliberately no fallback: if the shared library is missing, or no B200 is visible, every
compute entry point raises ``NativeError``.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libutmos_b200.so")

AF_NONE, AF_F64, AF_F32 = 0, 1, 2
F_NO_TRANSPOSE, F_STEP_KERNELS, F_FORCE_TRANSPOSE, F_NO_CLUSTER, F_NO_TAIL, F_DSMEM_GAINS, F_REF_TIES = 1, 2, 4, 8, 16, 32, 64
STOP_NONE, STOP_ZERO, STOP_ALL = 0, 1, 2
E_NOGPU = -3

# every symbol include/utmos_b200.h declares (tests check the library exports all of them)
SYMBOLS = ["utmos_last_error", "utmos_version", "utmos_device_count", "utmos_host_alloc", "utmos_host_free",
           "utmos_create", "utmos_destroy", "utmos_append_packed", "utmos_append_packed_device",
           "utmos_append_dense_u8", "utmos_append_dense_f32", "utmos_append_h5_chunks", "utmos_finalize", "utmos_select_begin",
           "utmos_append_packed2", "utmos_select_steps", "utmos_select_export", "utmos_select_import", "utmos_convert_gt", "utmos_convert_gt_ex", "utmos_convert_kernel_ms", "utmos_rows", "utmos_mgpu_layout", "utmos_mgpu_export", "utmos_mgpu_connect", "utmos_get_gains0", "utmos_set_gains0",
           "utmos_debug_gains", "utmos_debug_step_times", "utmos_debug_counters", "utmos_set_option", "utmos_info", "utmos_timings", "utmos_timer_start", "utmos_timer_stop",
           "utmos_lzf_decompress", "utmos_lzf_compress", "utmos_h5_encode_chunks", "utmos_gz_size", "utmos_gz_inflate", "utmos_vcf_parse_gt", "utmos_device_alloc", "utmos_device_free",
           "utmos_device_to_host", "utmos_synth_packed_device"]


class NativeError(RuntimeError):
    """A call into libutmos_b200.so failed."""

    def __init__(self, code, message):
        super().__init__(f"libutmos_b200 error {code}: {message}")
        self.code = code


_LIB = None


def lib():
    """Load the shared library (built in-tree by `python -m utmos_b200.build`)."""
    global _LIB  # pylint: disable=global-statement
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise NativeError(-100, f"{LIB_PATH} is missing: build it with `python -m utmos_b200.build` "
                                "(needs nvcc; there is no CPU fallback)")
    dll = ctypes.CDLL(LIB_PATH)
    i64, i32, u32, p = ctypes.c_int64, ctypes.c_int, ctypes.c_uint32, ctypes.c_void_p
    pp = ctypes.POINTER(ctypes.c_void_p)
    sig = {
        "utmos_last_error": (ctypes.c_char_p, []),
        "utmos_version": (ctypes.c_char_p, []),
        "utmos_device_count": (i32, [ctypes.POINTER(i32)]),
        "utmos_host_alloc": (i32, [pp, i64]),
        "utmos_host_free": (i32, [p]),
        "utmos_create": (i32, [pp, i32, i64, i64, i32, u32]),
        "utmos_destroy": (i32, [p]),
        "utmos_append_packed": (i32, [p, p, i64, i64, p]),
        "utmos_append_packed_device": (i32, [p, p, i64, i64, p]),
        "utmos_append_dense_u8": (i32, [p, p, i64]),
        "utmos_append_dense_f32": (i32, [p, p, i64]),
        "utmos_append_h5_chunks": (i32, [p, ctypes.c_char_p, i64, p, p, p, i64, i64, i32, i32, i32]),
        "utmos_finalize": (i32, [p, ctypes.POINTER(i64), p]),
        "utmos_select_begin": (i32, [p, p, p]),
        "utmos_select_steps": (i32, [p, i64, p, p, p, ctypes.POINTER(i64), ctypes.POINTER(i32)]),
        "utmos_append_packed2": (i32, [p, p, p, i64, i32, p]),
        "utmos_select_export": (i32, [p, p, p, i64, p, p, p, i64, ctypes.POINTER(i64), ctypes.POINTER(i64),
                                      ctypes.POINTER(ctypes.c_int)]),
        "utmos_select_import": (i32, [p, p, p, p, i64, p, p, p, i64, i64, ctypes.c_int]),
        "utmos_convert_gt": (i32, [i32, p, i64, i64, i64, p, p, ctypes.POINTER(i64), ctypes.POINTER(i64), p]),
        "utmos_convert_gt_ex": (i32, [i32, p, i64, i64, i64, p, p, ctypes.POINTER(i64), ctypes.POINTER(i64), p, i32]),
        "utmos_convert_kernel_ms": (i32, [ctypes.POINTER(ctypes.c_double)]),
        "utmos_debug_gains": (i32, [p, p, p]),
        "utmos_info": (i32, [p, p, i32]),
        "utmos_debug_step_times": (i32, [p, i64, i64, p]),
        "utmos_set_option": (i32, [p, i32, i64]),
        "utmos_debug_counters": (i32, [p, p]),
        "utmos_rows": (i32, [p, ctypes.POINTER(i64)]),
        "utmos_mgpu_layout": (i32, [p, i64, i64, i32]),
        "utmos_mgpu_export": (i32, [p, i32, i32, p]),
        "utmos_mgpu_connect": (i32, [p, p]),
        "utmos_get_gains0": (i32, [p, p, p, p]),
        "utmos_set_gains0": (i32, [p, p, p, p, i64]),
        "utmos_timings": (i32, [p, p, i32, i32]),
        "utmos_timer_start": (i32, [i32]),
        "utmos_timer_stop": (i32, [i32, ctypes.POINTER(ctypes.c_double)]),
        "utmos_lzf_decompress": (i64, [p, i64, p, i64]),
        "utmos_lzf_compress": (i64, [p, i64, p, i64]),
        "utmos_h5_encode_chunks": (i32, [p, i64, i64, i64, p, i64, p, p, p, i32]),
        "utmos_gz_size": (i64, [p, i64, ctypes.POINTER(i32)]),
        "utmos_gz_inflate": (i32, [p, i64, p, i64, ctypes.POINTER(i64), i32]),
        "utmos_vcf_parse_gt": (i32, [p, i64, i64, p, i64, ctypes.POINTER(i64), ctypes.POINTER(i64), i32, i32]),
        "utmos_device_alloc": (i32, [i32, pp, i64]),
        "utmos_device_free": (i32, [i32, p]),
        "utmos_device_to_host": (i32, [i32, p, p, i64]),
        "utmos_synth_packed_device": (i32, [i32, ctypes.c_uint64, i64, i64, i64, p, p, i64, p, p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(dll, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = dll
    return dll


def check(code):
    """Raise NativeError for a non-zero return code."""
    if code != 0:
        raise NativeError(code, lib().utmos_last_error().decode("utf-8", "replace"))


def timer_start(device=0):
    """Device stopwatch: synchronise the device and record the start CUDA event."""
    check(lib().utmos_timer_start(int(device)))


def timer_stop(device=0):
    """Synchronise the device, record the stop event; milliseconds between the two events."""
    ms = ctypes.c_double(0.0)
    check(lib().utmos_timer_stop(int(device), ctypes.byref(ms)))
    return ms.value


def device_count():
    """Number of visible CUDA devices (0 on a CPU box)."""
    n = ctypes.c_int(0)
    check(lib().utmos_device_count(ctypes.byref(n)))
    return n.value


def _ptr(arr):
    return None if arr is None else arr.ctypes.data_as(ctypes.c_void_p)


class PinnedBuffer:
    """Page-locked host memory exposed as a numpy uint8 array (zero-staging H2D copies)."""

    def __init__(self, nbytes):
        self._ptr = ctypes.c_void_p()
        check(lib().utmos_host_alloc(ctypes.byref(self._ptr), int(nbytes)))
        self.nbytes = int(nbytes)
        buf = (ctypes.c_uint8 * max(self.nbytes, 1)).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=np.uint8, count=self.nbytes)

    def close(self):
        if self._ptr is not None and self._ptr.value:
            self.array = None
            lib().utmos_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # pylint: disable=broad-except
            pass


class DeviceMatrix:
    """Bit-packed presence matrix resident in HBM plus the greedy-selection state machine.

    Stands in for the dense ``data['data']`` ndarray / h5py dataset of the reference: exposes ``.shape``
    ``(num_vars, num_samples)`` and ``.dtype`` (bool, float64 or float32), which is all ``select_main`` and
    ``run_selection`` look at (utmos/select.py:153, :429-433).
    """

    def __init__(self, n_samples, af_mode=AF_NONE, rows_hint=0, device=0, flags=0):
        self._ctx = ctypes.c_void_p()
        self.n_samples = int(n_samples)
        self.af_mode = af_mode
        self.num_vars = None
        self.var_count = None
        env_flags = int(os.environ.get("UTMOS_B200_FLAGS", "0"))  # test / debugging override of the kernel flavour
        if env_flags:
            flags = (flags & ~F_REF_TIES) | env_flags             # an explicit flavour wins over the tie-replay flavour
        check(lib().utmos_create(ctypes.byref(self._ctx), device, self.n_samples, int(rows_hint), af_mode, flags))
        for item in filter(None, os.environ.get("UTMOS_B200_OPTIONS", "").split(",")):   # "10=1,3=4096": A/B runs of the tuning options
            option, value = item.split("=")
            self.set_option(int(option), int(value))

    # -- ingestion ---------------------------------------------------------------------------------
    def append_packed(self, gt, af=None):
        """One .jl part: uint8 [V, >=ceil(S/8)] MSB-first rows (+ float64 AF per row)."""
        gt = np.ascontiguousarray(gt, dtype=np.uint8)
        if gt.ndim != 2:
            raise ValueError("GT must be 2-D")
        if af is not None:
            af = np.ascontiguousarray(np.asarray(af, dtype=np.float64).reshape(-1))
            if len(af) != gt.shape[0]:
                raise ValueError("AF length does not match GT rows")
        if self.af_mode == AF_NONE:
            af = None
        check(lib().utmos_append_packed(self._ctx, _ptr(gt), gt.shape[0], gt.shape[1], _ptr(af)))

    def append_packed_device(self, d_rows, n_rows, pitch, d_af=None):
        """Rows already resident in HBM (raw device pointers as ints)."""
        check(lib().utmos_append_packed_device(self._ctx, ctypes.c_void_p(d_rows), n_rows, pitch,
                                               ctypes.c_void_p(d_af) if d_af else None))

    def append_packed2(self, gt2, af=None):
        """One `.jl` v2 part (utmos_b200/jl2.py): row-compressed rows, decoded on the GPU (utmos_append_packed2)."""
        if "offsets" not in gt2:
            from utmos_b200 import jl2  # pylint: disable=import-outside-toplevel
            gt2 = jl2.with_offsets(gt2)
        payload = np.ascontiguousarray(gt2["payload"], dtype=np.uint8)
        offsets = np.ascontiguousarray(gt2["offsets"], dtype=np.uint64)
        if int(gt2["n_samples"]) != self.n_samples:
            raise ValueError("GT2 part has another sample count")
        n = len(offsets) - 1
        if af is not None:
            af = np.ascontiguousarray(np.asarray(af, dtype=np.float64).reshape(-1))
            if len(af) != n:
                raise ValueError("AF length does not match the rows")
        if len(payload) == 0:
            payload = np.zeros(1, dtype=np.uint8)
        check(lib().utmos_append_packed2(self._ctx, _ptr(payload), _ptr(offsets), n, int(gt2["idx_bytes"]), _ptr(af)))

    def append_dense(self, chunk):
        """One hdf5 chunk: bool/uint8 [n, S] or float32 [n, S] (GT*AF)."""
        if chunk.dtype == np.float32:
            chunk = np.ascontiguousarray(chunk)
            check(lib().utmos_append_dense_f32(self._ctx, _ptr(chunk), chunk.shape[0]))
        else:
            chunk = np.ascontiguousarray(chunk).view(np.uint8)
            check(lib().utmos_append_dense_u8(self._ctx, _ptr(chunk), chunk.shape[0]))

    def append_h5_chunks(self, path, addr, nbytes, fmask, rows_per_chunk, total_rows, is_f32, lzf=True, threads=0):
        """Chunks of an hdf5 'data' dataset (byte ranges in row order), read and LZF-decoded by native host threads."""
        addr = np.ascontiguousarray(addr, dtype=np.int64)
        nbytes = np.ascontiguousarray(nbytes, dtype=np.int64)
        fmask = np.ascontiguousarray(fmask, dtype=np.uint32)
        check(lib().utmos_append_h5_chunks(self._ctx, os.fsencode(path), len(addr), _ptr(addr), _ptr(nbytes), _ptr(fmask),
                                           int(rows_per_chunk), int(total_rows), 1 if is_f32 else 0, 1 if lzf else 0,
                                           int(threads)))

    def finalize(self):
        """Close ingestion; returns var_count (int64[S])."""
        nv = ctypes.c_int64(0)
        vc = np.zeros(self.n_samples, dtype=np.int64)
        check(lib().utmos_finalize(self._ctx, ctypes.byref(nv), _ptr(vc)))
        self.num_vars = nv.value
        self.var_count = vc
        return vc

    # -- multi-GPU plumbing (see utmos_b200/distributed.py) ------------------------------------------
    def rows(self):
        """Informative rows kept so far (synchronises the ingest stream)."""
        n = ctypes.c_int64(0)
        check(lib().utmos_rows(self._ctx, ctypes.byref(n)))
        return n.value

    def mgpu_layout(self, row_base, merged_rows, allow_tail=True):
        check(lib().utmos_mgpu_layout(self._ctx, int(row_base), int(merged_rows), 1 if allow_tail else 0))

    def mgpu_export(self, rank, world):
        handle = np.zeros(64, dtype=np.uint8)
        check(lib().utmos_mgpu_export(self._ctx, rank, world, _ptr(handle)))
        return handle

    def mgpu_connect(self, handles):
        handles = np.ascontiguousarray(handles, dtype=np.uint8)
        check(lib().utmos_mgpu_connect(self._ctx, _ptr(handles)))

    def get_gains0(self):
        cnt = np.zeros(self.n_samples, dtype=np.uint32)
        lo = np.zeros(self.n_samples, dtype=np.uint64)
        hi = np.zeros(self.n_samples, dtype=np.uint64)
        check(lib().utmos_get_gains0(self._ctx, _ptr(cnt), _ptr(lo), _ptr(hi)))
        return cnt, lo, hi

    def set_gains0(self, cnt, lo, hi, global_rows):
        cnt = np.ascontiguousarray(cnt, dtype=np.uint32)
        lo = np.ascontiguousarray(lo, dtype=np.uint64)
        hi = np.ascontiguousarray(hi, dtype=np.uint64)
        check(lib().utmos_set_gains0(self._ctx, _ptr(cnt), _ptr(lo), _ptr(hi), int(global_rows)))

    # -- what the reference looks at ---------------------------------------------------------------
    @property
    def shape(self):
        return (self.num_vars, self.n_samples)

    @property
    def dtype(self):
        return {AF_NONE: np.dtype(bool), AF_F64: np.dtype(np.float64), AF_F32: np.dtype(np.float32)}[self.af_mode]

    # -- selection ---------------------------------------------------------------------------------
    def begin(self, mask, weights=None):
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        if weights is not None:
            weights = np.ascontiguousarray(weights, dtype=np.float64)
        check(lib().utmos_select_begin(self._ctx, _ptr(mask), _ptr(weights)))

    def steps(self, max_steps):
        """Up to max_steps greedy steps -> (idx int64[n], new int64[n], score float64[n], stop)."""
        cap = max(1, min(int(max_steps), self.n_samples))
        idx = np.zeros(cap, dtype=np.int64)
        new = np.zeros(cap, dtype=np.int64)
        score = np.zeros(cap, dtype=np.float64)
        n = ctypes.c_int64(0)
        stop = ctypes.c_int(0)
        check(lib().utmos_select_steps(self._ctx, int(max_steps), _ptr(idx), _ptr(new), _ptr(score),
                                       ctypes.byref(n), ctypes.byref(stop)))
        return idx[:n.value], new[:n.value], score[:n.value], stop.value

    # -- resume --------------------------------------------------------------------------------------
    def export_state(self):
        """State of the selection in progress (utmos_select_export): dict of mask, live, idx, new, score, tot, stop."""
        words = self.info()["live_words"]
        mask = np.zeros(self.n_samples, dtype=np.uint8)
        live = np.zeros(words, dtype=np.uint32)
        idx = np.zeros(self.n_samples, dtype=np.int64)
        new = np.zeros(self.n_samples, dtype=np.int64)
        score = np.zeros(self.n_samples, dtype=np.float64)
        n, tot, stop = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int(0)
        check(lib().utmos_select_export(self._ctx, _ptr(mask), _ptr(live), words, _ptr(idx), _ptr(new), _ptr(score),
                                        self.n_samples, ctypes.byref(n), ctypes.byref(tot), ctypes.byref(stop)))
        return {"mask": mask, "live": live, "idx": idx[:n.value].copy(), "new": new[:n.value].copy(),
                "score": score[:n.value].copy(), "tot": int(tot.value), "stop": int(stop.value)}

    def import_state(self, state, weights=None):
        """Continue a selection from ``export_state()`` output on a context holding the same matrix."""
        mask = np.ascontiguousarray(state["mask"], dtype=np.uint8)
        live = np.ascontiguousarray(state["live"], dtype=np.uint32)
        idx = np.ascontiguousarray(state["idx"], dtype=np.int64)
        new = np.ascontiguousarray(state["new"], dtype=np.int64)
        score = np.ascontiguousarray(state["score"], dtype=np.float64)
        if weights is not None:
            weights = np.ascontiguousarray(weights, dtype=np.float64)
        check(lib().utmos_select_import(self._ctx, _ptr(mask), _ptr(weights), _ptr(live), len(live), _ptr(idx), _ptr(new),
                                        _ptr(score), len(idx), int(state["tot"]), int(state["stop"])))

    # -- introspection -----------------------------------------------------------------------------
    def gains(self):
        cnt = np.zeros(self.n_samples, dtype=np.int64)
        score = np.zeros(self.n_samples, dtype=np.float64)
        check(lib().utmos_debug_gains(self._ctx, _ptr(cnt), _ptr(score)))
        return cnt, score

    def step_times(self, first, n):
        """%globaltimer ns of the picks first..first+n-1 (profiling aid)."""
        out = np.zeros(n, dtype=np.int64)
        check(lib().utmos_debug_step_times(self._ctx, first, n, _ptr(out)))
        return out

    def counters(self):
        out = np.zeros(16, dtype=np.int64)
        check(lib().utmos_debug_counters(self._ctx, _ptr(out)))
        return out

    def set_option(self, option, value):
        """utmos_set_option (include/utmos_b200.h UTMOS_OPT_*): 1 regain rows, 2 per-step timestamps, 3 tail hand-over rows,
        5 owner-computes-cluster -> single-SM rows, 10 entry-divided-cluster -> shared-memory tail rows, 11 list budget
        (before finalize)."""
        check(lib().utmos_set_option(self._ctx, int(option), int(value)))

    def set_regain_rows(self, rows):
        """Override the recompute-vs-subtract threshold (0 never recompute, -1 default)."""
        check(lib().utmos_set_option(self._ctx, 1, int(rows)))

    def info(self):
        arr = np.zeros(10, dtype=np.int64)
        check(lib().utmos_info(self._ctx, _ptr(arr), 10))
        keys = ["num_vars", "row_pitch_bytes", "has_sample_major", "device_bytes", "fixed_scale", "af_inexact",
                "kernel_launches", "flavour", "live_words", "ref_ties"]
        return dict(zip(keys, (int(x) for x in arr)))

    def timings(self, reset=False):
        arr = np.zeros(8, dtype=np.float64)
        check(lib().utmos_timings(self._ctx, _ptr(arr), 8, 1 if reset else 0))
        return dict(zip(["h2d_ms", "ingest_ms", "transpose_ms", "gain_ms", "select_ms", "head_ms", "handover_ms",
                         "tail_ms"], (float(x) for x in arr)))

    def close(self):
        if self._ctx is not None and self._ctx.value:
            lib().utmos_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # pylint: disable=broad-except
            pass


CVT_DROP_SINGLETONS = 1


def convert_gt(gt, device=0, drop_singletons=False):
    """utmos/convert.py:57-87 on the GPU.  gt: int8 [V, S, ploidy].  Returns (packed, af(V,1), num_het, num_hom,
    singleton flags).  ``drop_singletons`` (--no-singleton, convert.py:58-62): the het / hom totals leave the flagged
    rows out; the caller drops their packed rows and AF (one kernel launch instead of two)."""
    gt = np.ascontiguousarray(gt, dtype=np.int8)
    if gt.ndim != 3:
        raise ValueError("GT tensor must be [variants, samples, ploidy]")
    n_vars, n_samples, ploidy = gt.shape
    packed = np.zeros((n_vars, (n_samples + 7) // 8), dtype=np.uint8)
    af = np.zeros(n_vars, dtype=np.float64)
    single = np.zeros(n_vars, dtype=np.uint8)
    het = ctypes.c_int64(0)
    hom = ctypes.c_int64(0)
    check(lib().utmos_convert_gt_ex(device, _ptr(gt), n_vars, n_samples, ploidy, _ptr(packed), _ptr(af),
                                    ctypes.byref(het), ctypes.byref(hom), _ptr(single),
                                    CVT_DROP_SINGLETONS if drop_singletons else 0))
    return packed, af.reshape(-1, 1), het.value, hom.value, single.astype(bool)


def convert_kernel_ms():
    """CUDA-event milliseconds of the K1 kernel launches of the last convert_gt call."""
    ms = ctypes.c_double(0.0)
    check(lib().utmos_convert_kernel_ms(ctypes.byref(ms)))
    return ms.value


def gz_inflate(data, threads=0):
    """gzip / BGZF bytes -> uint8 array (BGZF blocks are inflated in parallel by native host threads)."""
    src = np.frombuffer(data, dtype=np.uint8)
    is_bgzf = ctypes.c_int(0)
    size = lib().utmos_gz_size(_ptr(src), len(src), ctypes.byref(is_bgzf))
    if size < 0:
        raise ValueError("not a gzip stream")
    if is_bgzf.value:
        dst = np.empty(max(size, 1), dtype=np.uint8)
        got = ctypes.c_int64(0)
        check(lib().utmos_gz_inflate(_ptr(src), len(src), _ptr(dst), len(dst), ctypes.byref(got), int(threads)))
        return dst[:got.value]
    import zlib  # pylint: disable=import-outside-toplevel
    # plain gzip: the trailer's ISIZE is only the size modulo 2^32 and members may be concatenated -> stream it
    out, dec, view = [], zlib.decompressobj(31), memoryview(data)
    while len(view):
        out.append(dec.decompress(view))
        view = memoryview(dec.unused_data)
        if dec.eof and len(view):
            dec = zlib.decompressobj(31)
        elif not dec.eof:
            break
    return np.frombuffer(b"".join(out), dtype=np.uint8)


def vcf_parse_gt(text, offset, n_samples, max_variants, final=True, threads=0):
    """Tokenise up to max_variants data lines of VCF text (uint8 array) starting at `offset` ->
    (int8 GT[n, S, 2], bytes consumed)."""
    text = np.ascontiguousarray(text, dtype=np.uint8)
    gt = np.empty((max(int(max_variants), 1), int(n_samples), 2), dtype=np.int8)
    n = ctypes.c_int64(0)
    used = ctypes.c_int64(0)
    view = text[offset:]
    check(lib().utmos_vcf_parse_gt(_ptr(view), len(view), int(n_samples), _ptr(gt), int(max_variants), ctypes.byref(n),
                                   ctypes.byref(used), 1 if final else 0, int(threads)))
    return gt[:n.value], used.value


def lzf_decompress(data, out_len):
    """liblzf-format block -> bytes of exactly out_len (hdf5 filter 32000; host code in the native lib)."""
    src = np.frombuffer(data, dtype=np.uint8)
    dst = np.empty(out_len, dtype=np.uint8)
    got = lib().utmos_lzf_decompress(_ptr(src), len(src), _ptr(dst), out_len)
    if got != out_len:
        raise ValueError(f"lzf: expected {out_len} bytes, decoded {got}")
    return dst


def lzf_compress(data):
    """bytes -> liblzf-format block, or None when it does not shrink (stored raw, filter_mask bit 0)."""
    src = np.frombuffer(data, dtype=np.uint8)
    dst = np.empty(max(len(src) - 1, 1), dtype=np.uint8)
    got = lib().utmos_lzf_compress(_ptr(src), len(src), _ptr(dst), len(dst))
    if got <= 0:
        return None
    return dst[:got].tobytes()


def h5_encode_chunks(gt_packed, n_samples, af, chunk_rows, threads=0):
    """Packed .jl rows -> the LZF chunks of a --lowmem 'data' dataset (bool bytes, or float32 GT*AF when af is given),
    unpacked and compressed by native host threads.  Returns [(bytes, filter_mask)] in chunk order."""
    gt_packed = np.ascontiguousarray(gt_packed, dtype=np.uint8)
    n_rows, pitch = gt_packed.shape
    if af is not None:
        af = np.ascontiguousarray(np.asarray(af, dtype=np.float64).reshape(-1))
    n_chunks = (n_rows + chunk_rows - 1) // chunk_rows
    chunk_nbytes = int(chunk_rows) * int(n_samples) * (4 if af is not None else 1)
    dst = np.empty(max(n_chunks * chunk_nbytes, 1), dtype=np.uint8)
    sizes = np.zeros(max(n_chunks, 1), dtype=np.int64)
    masks = np.zeros(max(n_chunks, 1), dtype=np.uint32)
    if n_chunks:
        check(lib().utmos_h5_encode_chunks(_ptr(gt_packed), n_rows, pitch, int(n_samples), _ptr(af), int(chunk_rows), _ptr(dst),
                                           _ptr(sizes), _ptr(masks), int(threads)))
    return [(dst[i * chunk_nbytes:i * chunk_nbytes + int(sizes[i])].tobytes(), int(masks[i])) for i in range(n_chunks)]


class DeviceBuffer:
    """Raw cudaMalloc'd buffer (benchmark inputs resident in HBM)."""

    def __init__(self, nbytes, device=0):
        self.device = device
        self.nbytes = int(nbytes)
        self._ptr = ctypes.c_void_p()
        check(lib().utmos_device_alloc(device, ctypes.byref(self._ptr), self.nbytes))

    @property
    def ptr(self):
        return self._ptr.value

    def to_host(self, out=None):
        if out is None:
            out = np.empty(self.nbytes, dtype=np.uint8)
        check(lib().utmos_device_to_host(self.device, _ptr(out), self._ptr, self.nbytes))
        return out

    def close(self):
        if self._ptr is not None and self._ptr.value:
            lib().utmos_device_free(self.device, self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # pylint: disable=broad-except
            pass


def synth_packed_device(seed, row0, n_rows, n_samples, cdf_thr, p_thr, d_rows, d_af, device=0):
    """Fill device buffers with synthetic cohort rows (see utmos_b200/synth.py for the tables / mirror)."""
    cdf_thr = np.ascontiguousarray(cdf_thr, dtype=np.uint64)
    p_thr = np.ascontiguousarray(p_thr, dtype=np.uint32)
    check(lib().utmos_synth_packed_device(device, seed, row0, n_rows, n_samples, _ptr(cdf_thr), _ptr(p_thr),
                                          len(cdf_thr) - 1, ctypes.c_void_p(d_rows.ptr), ctypes.c_void_p(d_af.ptr)))
