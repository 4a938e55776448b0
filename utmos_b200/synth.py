"""Deterministic synthetic cohorts in the .jl layout.

Allele count k_v of variant v follows a truncated power law P(k) ~ k^-1.4 on [1, 2S-1]; AF_v = k_v/(2S);
carrier probability p_v = 1-(1-AF_v)^2 (Hardy-Weinberg); bit(v, s) ~ Bernoulli(p_v) from a counter-based
hash of (seed, v, s) (splitmix64 finaliser); at least one carrier per row, so every row is informative.

The cohort is generated on the GPU (csrc/synth.cu) straight into HBM; ``mirror_rows`` is the bit-identical
NumPy restatement used by the tests at reduced shapes.  Both consume the same integer tables built here.
"""
import numpy as np

from utmos_b200 import _native

M1 = np.uint64(0xBF58476D1CE4E5B9)
M2 = np.uint64(0x94D049BB133111EB)
GV = np.uint64(0x9E3779B97F4A7C15)
GS = np.uint64(0xD1B54A32D192ED03)
ALPHA = 1.4


def tables(n_samples, alpha=ALPHA):
    """(cdf_thr uint64[kmax+1], p_thr uint32[kmax+1]); entry 0 unused."""
    kmax = 2 * n_samples - 1
    k = np.arange(1, kmax + 1, dtype=np.float64)
    w = k ** (-alpha)
    cdf = np.cumsum(w) / w.sum()
    cdf_thr = np.zeros(kmax + 1, dtype=np.uint64)
    scaled = np.minimum(cdf * 18446744073709551616.0, 18446744073709549568.0)     # largest double < 2^64
    cdf_thr[1:] = scaled.astype(np.uint64)
    cdf_thr[kmax] = np.uint64(0xFFFFFFFFFFFFFFFF)
    af = k / (2.0 * n_samples)
    p = 1.0 - (1.0 - af) ** 2
    p_thr = np.zeros(kmax + 1, dtype=np.uint32)
    p_thr[1:] = np.minimum(np.floor(p * 4294967296.0), 4294967295.0).astype(np.uint32)
    return cdf_thr, p_thr


def _mix64(z):
    z = (z ^ (z >> np.uint64(30))) * M1
    z = (z ^ (z >> np.uint64(27))) * M2
    return z ^ (z >> np.uint64(31))


def _cell_hash(seed, v, s):
    with np.errstate(over="ignore"):
        return _mix64(np.uint64(seed) + v.astype(np.uint64) * GV + (s.astype(np.uint64) + np.uint64(1)) * GS)


def mirror_rows(seed, row0, n_rows, n_samples, tabs=None):
    """NumPy restatement of synth_rows_kernel: (uint8 [n_rows, ceil(S/8)] MSB-first, float64 AF[n_rows])."""
    cdf_thr, p_thr = tabs if tabs is not None else tables(n_samples)
    kmax = len(cdf_thr) - 1
    v = np.arange(row0, row0 + n_rows, dtype=np.uint64)
    u = _cell_hash(seed, v, np.full(n_rows, 0xFFFFFFFF, dtype=np.uint64))
    k = np.searchsorted(cdf_thr[1:], u, side="right") + 1          # smallest k with u < cdf_thr[k]
    k = np.minimum(k, kmax)
    thr = p_thr[k].astype(np.uint64)
    s = np.arange(n_samples, dtype=np.uint64)
    h = _cell_hash(seed, v[:, None], s[None, :]) >> np.uint64(32)
    bits = h < thr[:, None]
    empty = ~bits.any(axis=1)
    if empty.any():
        pos = _cell_hash(seed, v[empty], np.full(int(empty.sum()), 0xFFFFFFFE, dtype=np.uint64)) % np.uint64(n_samples)
        bits[np.nonzero(empty)[0], pos.astype(np.int64)] = True
    return np.packbits(bits, axis=1), k / (2.0 * n_samples)


class DeviceCohort:
    """A synthetic cohort generated in HBM: .jl-layout rows + AF as raw device buffers."""

    def __init__(self, seed, n_vars, n_samples, device=0, row0=0):
        self.seed, self.n_vars, self.n_samples, self.device = seed, n_vars, n_samples, device
        self.pitch = (n_samples + 7) // 8
        self.rows = _native.DeviceBuffer(n_vars * self.pitch, device)
        self.af = _native.DeviceBuffer(n_vars * 8, device)
        cdf_thr, p_thr = tables(n_samples)
        _native.synth_packed_device(seed, row0, n_vars, n_samples, cdf_thr, p_thr, self.rows, self.af, device)

    def to_host(self, rows_out=None, af_out=None):
        """Copy to host (optionally into caller buffers, e.g. pinned): (uint8 [V, pitch], float64 [V])."""
        if rows_out is None:
            rows_out = np.empty(self.n_vars * self.pitch, dtype=np.uint8)
        if af_out is None:
            af_out = np.empty(self.n_vars, dtype=np.float64)
        self.rows.to_host(rows_out.reshape(-1).view(np.uint8))
        self.af.to_host(af_out.view(np.uint8))
        return rows_out.reshape(self.n_vars, self.pitch), af_out

    def close(self):
        self.rows.close()
        self.af.close()


def sample_names(n_samples):
    return np.array([f"S{i:07d}" for i in range(n_samples)])


def synthetic_weights(n_samples, seed=0):
    """1.0 everywhere except 1% of the samples, which get integers 2..10 (mirrors weights.txt)."""
    rng = np.random.default_rng([seed, 0xBEEF])
    w = np.ones(n_samples)
    pick = rng.choice(n_samples, max(1, n_samples // 100), replace=False)
    w[pick] = rng.integers(2, 11, len(pick))
    return w
