"""NumPy restatement of the utmos hot path.  TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` leg
may import this module, and only as the checker or the CPU baseline.  Nothing under ``utmos_b200/``
imports it; the shipped path has no CPU fallback.

Parity status: PINNED (tests/test_oracle.py): reproduces every ``.jl`` / hdf5 answer key of the reference
and the outputs of the unmodified reference ``select.py`` recorded in ``tests/golden/*.json``.

Two flavours live here:

``DenseOracle``
    operates on the same dense ``[V', S]`` bool / float matrix the reference builds and performs the same
    NumPy operations row by row, in the same order (utmos/select.py:24-53, :69-112), so that float64
    scores are bit-identical and the CPU cost has the reference's shape (one interpreted iteration per
    variant row per step).  This is what ``bench.py`` times as the CPU baseline (``kind: "port"``).

``greedy_c``
    ctypes binding to ``oracle/greedy_oracle.c`` (same algorithm on packed bits, plain C) for shapes
    where a dense float64 matrix does not fit or Python is too slow.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")

STOP_COUNT, STOP_ZERO, STOP_ALL = 0, 1, 2


# --------------------------------------------------------------------------------------------------
# load stage: utmos/select.py:272-284, :314-320
# --------------------------------------------------------------------------------------------------
def unpack_and_filter(packed, n_samples, af=None):
    """utmos/select.py:275-280.  Returns (bool matrix of informative rows, their AF, keep flags)."""
    dense = np.unpackbits(np.ascontiguousarray(packed), axis=1, count=n_samples).astype(bool)
    informative = dense.any(axis=1)
    kept_af = None if af is None else np.asarray(af)[informative]
    return dense[informative], kept_af, informative


def load_parts(parts, n_samples, with_af=False, float32_af=False):
    """Fold several ``.jl`` payloads ({'GT','AF'}) the way load_files does (utmos/select.py:262-320).

    Returns ``(matrix, var_count)``; ``matrix`` is bool, or float64 ``GT*AF`` when ``with_af``
    (select.py:317-320), or float32 when ``float32_af`` (the hdf5 flavour, select.py:218-223).
    """
    mats, afs = [], []
    var_count = np.zeros(n_samples, dtype=np.int64)
    for part in parts:
        dense, kept_af, _ = unpack_and_filter(part["GT"], n_samples, part["AF"])
        mats.append(dense)
        afs.append(kept_af)
        var_count += dense.sum(axis=0)
    matrix = np.concatenate(mats) if len(mats) > 1 else mats[0]
    if with_af or float32_af:
        af_col = np.concatenate(afs) if len(afs) > 1 else afs[0]
        matrix = matrix * af_col.reshape(-1, 1)
        if float32_af:
            matrix = matrix.astype(np.float32)
    return matrix, var_count


# --------------------------------------------------------------------------------------------------
# selection setup: utmos/select.py:153-187
# --------------------------------------------------------------------------------------------------
def resolve_count(select_count, n_samples):
    """utmos/select.py:157-159."""
    if select_count < 0:
        return n_samples
    if select_count < 1:
        return max(1, int(n_samples * select_count))
    return max(1, int(select_count))


def build_mask(names, subset=None, exclude=None):
    """utmos/select.py:168-175: 1 selectable, 2 excluded; subset first, then exclude."""
    names = np.asarray(names).astype(str)
    mask = np.ones(len(names), dtype=np.uint8)
    if subset:
        mask = np.where(np.isin(names, subset), 1, 2).astype(np.uint8)
    if exclude:
        mask = np.where(np.isin(names, exclude), 2, mask).astype(np.uint8)
    return mask


def build_weights(names, weight_map):
    """utmos/select.py:181-187: ones, overwritten for names present in the weights table."""
    if weight_map is None:
        return None
    out = np.ones(len(names), dtype=np.float64)
    for pos, name in enumerate(np.asarray(names).astype(str)):
        if name in weight_map:
            out[pos] = weight_map[name]
    return out


# --------------------------------------------------------------------------------------------------
# the greedy loop: utmos/select.py:24-53 and :69-112
# --------------------------------------------------------------------------------------------------
class DenseOracle:
    """Row-by-row NumPy greedy max coverage with the reference's exact operation order."""

    def __init__(self, matrix, mask, weights=None):
        self.matrix = matrix
        self.mask = np.array(mask, dtype=np.uint8, copy=True)
        self.weights = None if weights is None else np.asarray(weights, dtype=np.float64)
        self.num_vars, self.num_samples = matrix.shape
        self.tot_captured = 0

    def score_once(self):
        """One pass of utmos/select.py:24-53.  Returns (best, new_count, best_score) or None when the
        best masked+weighted score is zero."""
        totals = np.zeros(self.num_samples)
        hits = np.zeros(self.num_samples, dtype="int")
        chosen = np.where(self.mask == 0)
        for row in self.matrix:                       # select.py:37
            if row[chosen].any():                     # select.py:38  covered by a selected sample
                continue
            totals += row                             # select.py:40
            hits += (row != 0).astype("int")          # select.py:41
        totals[self.mask != 1] = 0                    # select.py:43
        if self.weights is not None:
            totals *= self.weights                    # select.py:47
        best = int(np.argmax(totals))                 # select.py:48 (first maximum)
        if totals[best] == 0:                         # select.py:51
            return None
        return best, int(hits[best]), float(totals[best])

    def run(self, max_steps):
        """utmos/select.py:91-112.  Returns (idx[], new[], score[], stop_reason)."""
        idx, new, score = [], [], []
        reason = STOP_COUNT
        for _ in range(max_steps):
            got = self.score_once()
            if got is None:
                reason = STOP_ZERO
                break
            best, fresh, top = got
            self.tot_captured += fresh
            self.mask[best] = 0
            idx.append(best)
            new.append(fresh)
            score.append(top)
            if self.tot_captured >= self.num_vars:
                reason = STOP_ALL
                break
        return (np.array(idx, dtype=np.int64), np.array(new, dtype=np.int64),
                np.array(score, dtype=np.float64), reason)


def report_rows(names, var_count, idx, new, num_vars):
    """The five report columns, built with the reference's expressions (utmos/select.py:97-108)."""
    rows = []
    tot = 0
    for best, fresh in zip(idx, new):
        tot = tot + np.int64(fresh)
        rows.append([str(names[best]), int(var_count[best]), int(fresh), int(tot), round(tot / num_vars, 4)])
    return rows


def format_report(rows):
    """utmos/select.py:441-445."""
    out = ["sample\tvar_count\tnew_count\ttot_captured\tpct_captured\n"]
    for row in rows:
        out.append("\t".join(str(_) for _ in row) + "\n")
    return "".join(out)


# --------------------------------------------------------------------------------------------------
# C oracle binding
# --------------------------------------------------------------------------------------------------
def build_c_oracle(force=False):
    """Compile oracle/greedy_oracle.c -> oracle/liboracle.so (gcc, a couple of seconds)."""
    src = os.path.join(HERE, "greedy_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", LIB_PATH, src, "-lm"])
    return LIB_PATH


_LIB = None


def c_lib():
    """Load (building if necessary) the C oracle."""
    global _LIB  # pylint: disable=global-statement
    if _LIB is None:
        lib = ctypes.CDLL(build_c_oracle())
        i64, p = ctypes.c_int64, ctypes.c_void_p
        lib.oracle_filter_rows.restype = i64
        lib.oracle_filter_rows.argtypes = [p, i64, i64, p, p]
        lib.oracle_greedy_select.restype = i64
        lib.oracle_greedy_select.argtypes = [p, i64, i64, p, p, p, i64, ctypes.c_int, p, p, p, p]
        lib.oracle_convert_gt.restype = None
        lib.oracle_convert_gt.argtypes = [p, i64, i64, i64, p, p, p, p, p]
        lib.oracle_score_vector.restype = None
        lib.oracle_score_vector.argtypes = [p, i64, i64, p, p, p, ctypes.c_int, p, p]
        lib.oracle_fixed_scale.restype = ctypes.c_int
        lib.oracle_fixed_scale.argtypes = [i64]
        _LIB = lib
    return _LIB


def _ptr(arr):
    return None if arr is None else arr.ctypes.data_as(ctypes.c_void_p)


def filter_rows_c(packed, n_samples):
    """C version of utmos/select.py:275-284.  Returns (keep flags, var_count)."""
    packed = np.ascontiguousarray(packed, dtype=np.uint8)
    keep = np.zeros(packed.shape[0], dtype=np.uint8)
    var_count = np.zeros(n_samples, dtype=np.int64)
    c_lib().oracle_filter_rows(_ptr(packed), packed.shape[0], n_samples, _ptr(keep), _ptr(var_count))
    return keep.astype(bool), var_count


def greedy_c(packed, n_samples, mask, weights=None, af=None, max_steps=None, exact=False):
    """C greedy over already-filtered packed rows.  Returns (idx, new, score, stop_reason)."""
    packed = np.ascontiguousarray(packed, dtype=np.uint8)
    n_vars = packed.shape[0]
    assert packed.shape[1] == (n_samples + 7) // 8
    mask = np.array(mask, dtype=np.uint8, copy=True)
    weights = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    af = None if af is None else np.ascontiguousarray(np.asarray(af).reshape(-1), dtype=np.float64)
    if max_steps is None:
        max_steps = n_samples
    cap = max(1, min(max_steps, n_samples))
    idx = np.zeros(cap, dtype=np.int64)
    new = np.zeros(cap, dtype=np.int64)
    score = np.zeros(cap, dtype=np.float64)
    reason = ctypes.c_int(0)
    n_out = c_lib().oracle_greedy_select(_ptr(packed), n_vars, n_samples, _ptr(af), _ptr(mask), _ptr(weights),
                                         min(max_steps, cap), 1 if exact else 0, _ptr(idx), _ptr(new), _ptr(score),
                                         ctypes.byref(reason))
    return idx[:n_out], new[:n_out], score[:n_out], reason.value


def score_vector_c(packed, n_samples, mask, weights=None, af=None, exact=False):
    """One scoring pass (utmos/select.py:24-47) for the selection history encoded in ``mask`` (0 = used).

    Returns (scores float64[S] after mask and weights, counts int64[S])."""
    packed = np.ascontiguousarray(packed, dtype=np.uint8)
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    weights = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    af = None if af is None else np.ascontiguousarray(np.asarray(af).reshape(-1), dtype=np.float64)
    scores = np.zeros(n_samples, dtype=np.float64)
    counts = np.zeros(n_samples, dtype=np.int64)
    c_lib().oracle_score_vector(_ptr(packed), packed.shape[0], n_samples, _ptr(af), _ptr(mask), _ptr(weights),
                                1 if exact else 0, _ptr(scores), _ptr(counts))
    return scores, counts


def convert_gt_c(gt):
    """C version of utmos/convert.py:57-87 on an int8 [V, S, ploidy] tensor.

    Returns dict(GT packed uint8, AF float64 (V,1), stats, singleton flags)."""
    gt = np.ascontiguousarray(gt, dtype=np.int8)
    n_vars, n_samples, ploidy = gt.shape
    packed = np.zeros((n_vars, (n_samples + 7) // 8), dtype=np.uint8)
    af = np.zeros(n_vars, dtype=np.float64)
    het = ctypes.c_int64(0)
    hom = ctypes.c_int64(0)
    single = np.zeros(n_vars, dtype=np.uint8)
    c_lib().oracle_convert_gt(_ptr(gt), n_vars, n_samples, ploidy, _ptr(packed), _ptr(af), ctypes.byref(het),
                              ctypes.byref(hom), _ptr(single))
    return {"GT": packed, "AF": af.reshape(-1, 1), "stats": {"num_het": het.value, "num_hom": hom.value},
            "singleton": single.astype(bool)}


def fixed_scale(n_vars):
    """Fixed-point scale used by the exact mode / the CUDA --af path for ``n_vars`` informative rows."""
    return c_lib().oracle_fixed_scale(n_vars)
