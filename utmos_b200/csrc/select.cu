// select.cu -- the greedy maximum-coverage loop on the GPU.
//
// Reference semantics (utmos/select.py:24-53, :69-112): every step scores each sample by the sum over
// the rows no selected sample carries, takes np.argmax (first index among equals), stops on a zero
// best score, marks the sample used, accumulates tot_captured.  The reference recomputes the O(V*S)
// sum from scratch each step; here the per-sample gains are computed ONCE (K3) and then maintained
// incrementally: after a pick, the rows it newly covers are cleared from the live bitmask and only
// those rows are subtracted from every sample's gain (K5).  In count mode this is exact integer
// arithmetic; in the AF flavours the gains are exact fixed-point integers (two limbs) that are rounded
// to float64 once per comparison, so equal multisets give bit-equal scores in any order.
//
// Kernels:  K2b transpose_bits_kernel   variant-major -> sample-major copy (one time)
//           K3  colpop_kernel / row_gain_kernel   initial gains + var_count
//           K4  argmax (argmax_step_kernel, or phase A of the persistent kernel)
//           K5  cover (cover_step_kernel, or phase B of the persistent kernel)
#include <cooperative_groups.h>

#include <algorithm>
#include <stdlib.h>

#include "common.cuh"

namespace utmos {

namespace {

// ------------------------------------------------------------------------------------------------
// K2b: bit-matrix transpose.  CTA tile = 256 rows x 16 words (512 samples): rows are read as 64-byte
// pieces (streaming 128-bit loads; 36 KB of shared memory per CTA keeps 6 CTAs per SM in flight) into shared memory, every warp transposes 32x32 bit blocks in
// registers (5 butterfly stages of one shuffle + a masked merge each), and every sample gets its 256 row
// bits as one aligned 32-byte sector (two 128-bit stores).  Algorithmic bytes: read + write V*pitch.
// ------------------------------------------------------------------------------------------------
constexpr int kTRows = 256;
constexpr int kTWords = 16;
constexpr int kTInPitch = 17;       // words; odd pitch -> column reads are bank-conflict free
constexpr int kTOutPitch = 9;
constexpr size_t kTransposeSmem = ((size_t)kTRows * kTInPitch + (size_t)kTWords * 32 * kTOutPitch) * 4;

// lane l holds row l of a 32x32 bit block; afterwards lane j holds column j (bit i = old row i, bit j)
__device__ __forceinline__ uint32_t transpose32(uint32_t x, int lane)
{
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) {
        const uint32_t m = k == 16 ? 0x0000ffffu : k == 8 ? 0x00ff00ffu : k == 4 ? 0x0f0f0f0fu : k == 2 ? 0x33333333u : 0x55555555u;
        const uint32_t y = __shfl_xor_sync(0xffffffffu, x, k);
        x = (lane & k) ? ((x & ~m) | ((y & ~m) >> k)) : ((x & m) | ((y & m) << k));
    }
    return x;
}

// The 64-byte pieces are read with plain read-only loads.  A streaming (L1::no_allocate) load is looked up
// evict-first in L2, where the whole 128-byte line is fetched: the half the neighbouring column tile needs was gone
// again before that CTA asked for it, and DRAM read every line twice (ncu r1: 707 MB for a 353 MB matrix).
__global__ void __launch_bounds__(256) transpose_bits_kernel(const uint32_t *__restrict__ rows, long long V,
                                                             int pitchW, int S32, uint32_t *__restrict__ cols,
                                                             long long colPitchW, int col_tiles)
{
    extern __shared__ __align__(16) uint32_t t_smem[];
    uint32_t *s_in = t_smem;                                   // [kTRows][kTInPitch]
    uint32_t *s_out = t_smem + kTRows * kTInPitch;             // [1024][kTOutPitch]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // 1-D grid, column tile fastest: the CTAs that share the 128-byte lines of the same rows run back to back, so the
    // second 64-byte half of every line comes from L2 instead of DRAM
    const long long row_tile = blockIdx.x / col_tiles;
    const int col_tile = (int)(blockIdx.x - row_tile * col_tiles);
    const long long r0 = row_tile * kTRows;
    const int w0 = col_tile * kTWords;

    for (int i = threadIdx.x; i < kTRows * (kTWords / 4); i += 256) {
        const int row = i / (kTWords / 4), q = i % (kTWords / 4);
        const long long r = r0 + row;
        const int w = w0 + q * 4;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (r < V && w < pitchW) {
            const uint4 *src = reinterpret_cast<const uint4 *>(rows + r * pitchW + w);
            v = __ldg(src);
        }
        uint32_t *d = s_in + row * kTInPitch + q * 4;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    __syncthreads();
    // warp g transposes the 32 blocks of row group g (rows g*32 .. g*32+31)
    const int g = warp;
#pragma unroll 4
    for (int c = 0; c < kTWords; ++c) {
        const uint32_t x = transpose32(s_in[(g * 32 + lane) * kTInPitch + c], lane);
        s_out[(c * 32 + lane) * kTOutPitch + g] = x;
    }
    __syncthreads();
    for (int sl = threadIdx.x; sl < kTWords * 32; sl += 256) {
        const int s = w0 * 32 + sl;
        if (s >= S32) continue;
        const uint32_t *o = s_out + sl * kTOutPitch;
        uint4 *dst = reinterpret_cast<uint4 *>(cols + (long long)s * colPitchW + row_tile * (kTRows / 32));
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
}

// 32x32 bit block held by ONE thread, a[i] bit j -> a[j] bit i (LSB-first on both axes): 5 stages of 16 masked swaps
__device__ __forceinline__ void transpose32_regs(uint32_t (&a)[32])
{
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
        const uint32_t m = j == 16 ? 0x0000ffffu : j == 8 ? 0x00ff00ffu : j == 4 ? 0x0f0f0f0fu : j == 2 ? 0x33333333u : 0x55555555u;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            if ((k & j) == 0) {
                const uint32_t t = ((a[k] >> j) ^ a[k + j]) & m;
                a[k] ^= t << j;
                a[k + j] ^= t;
            }
        }
    }
}

// K2b, register flavour.  CTA tile = 256 rows x 32 words (1,024 samples): every row piece is one 128-byte run read by
// eight neighbouring threads, staged in shared memory, then thread (c, g) = (tid >> 3, tid & 7) takes the 32x32 bit
// block of word column c and row group g into 32 registers, transposes it there (5 stages of 16 masked swaps: no
// shuffles, no second staging buffer, one barrier per tile) and stores word g of samples c*32 .. c*32+31: for a given
// sample the eight lanes g = 0..7 write one aligned 32-byte sector.  Shared layout: row r at r*33 + (r >> 5)*4 words
// (the staging stores and the column reads are both bank-conflict free).
constexpr int kRRows = 256;
constexpr int kRWords = 32;
constexpr int kRPitch = 33;
constexpr size_t kTransposeRegSmem = ((size_t)kRRows * kRPitch + (kRRows / 32) * 4) * 4;

__global__ void __launch_bounds__(256, 5) transpose_bits_reg_kernel(const uint32_t *__restrict__ rows, long long V,
                                                                    int pitchW, int nWS, uint32_t *__restrict__ cols,
                                                                    long long colPitchW, int col_tiles)
{
    extern __shared__ __align__(16) uint32_t t_smem[];
    const long long row_tile = blockIdx.x / col_tiles;          // column tile fastest: the CTAs that share the
    const int col_tile = (int)(blockIdx.x - row_tile * col_tiles);   // 128-byte lines of a row run back to back (L2)
    const long long r0 = row_tile * kRRows;
    const int w0 = col_tile * kRWords;
#pragma unroll
    for (int it = 0; it < kRRows * (kRWords / 4) / 256; ++it) {
        const int i = it * 256 + threadIdx.x;
        const int row = i >> 3, q = i & 7;
        const long long r = r0 + row;
        const int w = w0 + q * 4;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (r < V && w < pitchW) v = __ldg(reinterpret_cast<const uint4 *>(rows + r * pitchW + w));
        uint32_t *d = t_smem + row * kRPitch + (row >> 5) * 4 + q * 4;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    __syncthreads();
    const int c = threadIdx.x >> 3, g = threadIdx.x & 7;
    if (w0 + c >= nWS) return;                                   // word columns past the last sample
    uint32_t a[32];
    const uint32_t *src = t_smem + (g * 32) * kRPitch + g * 4 + c;
#pragma unroll
    for (int i = 0; i < 32; ++i) a[i] = src[i * kRPitch];
    transpose32_regs(a);
    uint32_t *dst = cols + (long long)(w0 + c) * 32 * colPitchW + row_tile * (kRRows / 32) + g;
#pragma unroll
    for (int j = 0; j < 32; ++j) dst[(long long)j * colPitchW] = a[j];
}

// ------------------------------------------------------------------------------------------------
// fixed-point AF: q = AF * 2^scale split into two limbs of L bits (DESIGN.md "fixed-point AF")
// ------------------------------------------------------------------------------------------------
__global__ void fixed_af_kernel(const double *__restrict__ af, long long V, int af_mode, int L, int scale,
                                unsigned long long *__restrict__ q_lo, unsigned long long *__restrict__ q_hi,
                                SelState *st)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= V) return;
    double a = af[r];
    if (af_mode == UTMOS_AF_F32) a = (double)(float)a;      // utmos/select.py:218-223 stores float32
    unsigned long long lo = 0, hi = 0;
    if (!(a >= 0.0) || a > 1.0) {
        atomicAdd(&st->af_invalid, 1u);
    } else if (a > 0.0) {
        int e;
        const double m = frexp(a, &e);                        // a = m * 2^e, m in [0.5, 1)
        const unsigned long long mi = (unsigned long long)scalbn(m, 53);
        const int sh = e - 53 + scale;                        // value = mi * 2^sh, 53-bit mi
        unsigned long long x_lo, x_hi;
        if (sh >= 0) {
            x_lo = sh < 64 ? mi << sh : 0ull;
            x_hi = sh == 0 ? 0ull : (sh < 64 ? mi >> (64 - sh) : mi << (sh - 64));
        } else {
            const int d = -sh;
            if (d >= 64) { x_lo = 0; atomicAdd(&st->af_inexact, 1u); }
            else {
                if (mi & ((1ull << d) - 1ull)) atomicAdd(&st->af_inexact, 1u);
                x_lo = mi >> d;
            }
            x_hi = 0;
        }
        lo = x_lo & ((1ull << L) - 1ull);
        hi = (x_lo >> L) | (x_hi << (64 - L));
    }
    q_lo[r] = lo;
    q_hi[r] = hi;
}

// live[w] bit r set  <=>  row r exists and (count mode, or its AF is nonzero: an all-zero float row never
// scores and is never covered, utmos/select.py:38-41)
__global__ void live_init_kernel(uint32_t *live, long long colPitchW, long long V, const unsigned long long *q_lo,
                                 const unsigned long long *q_hi, int af)
{
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= colPitchW) return;
    uint32_t m = 0;
    for (int j = 0; j < 32; ++j) {
        const long long r = w * 32 + j;
        if (r < V && (!af || (q_lo[r] | q_hi[r]) != 0ull)) m |= 1u << j;
    }
    live[w] = m;
}

// ------------------------------------------------------------------------------------------------
// K3 (sample-major source): one CTA per sample, 128-bit streaming loads, popcount.
//   var_count[s] = popcount(col_s)            (utmos/select.py:281-284)
//   gain_cnt[s]  = popcount(col_s & live)     (step-0 counts of utmos/select.py:41)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colpop_kernel(const uint32_t *__restrict__ cols, const uint32_t *__restrict__ live,
                                                     long long colPitchW, int S, unsigned int *var_count,
                                                     unsigned int *gain_cnt)
{
    const int s = blockIdx.x;
    if (s >= S) return;
    const uint4 *col = reinterpret_cast<const uint4 *>(cols + (long long)s * colPitchW);
    const uint4 *lv = reinterpret_cast<const uint4 *>(live);
    const long long n4 = colPitchW / 4;
    unsigned int all = 0, alive = 0;
    for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
        const uint4 c = ld_stream_u128(col + i);
        const uint4 l = __ldg(lv + i);
        all += __popc(c.x) + __popc(c.y) + __popc(c.z) + __popc(c.w);
        alive += __popc(c.x & l.x) + __popc(c.y & l.y) + __popc(c.z & l.z) + __popc(c.w & l.w);
    }
    __shared__ unsigned int s_all[8], s_alive[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        all += __shfl_xor_sync(0xffffffffu, all, o);
        alive += __shfl_xor_sync(0xffffffffu, alive, o);
    }
    if ((threadIdx.x & 31) == 0) { s_all[threadIdx.x >> 5] = all; s_alive[threadIdx.x >> 5] = alive; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int a = 0, b = 0;
        for (int i = 0; i < 8; ++i) { a += s_all[i]; b += s_alive[i]; }
        if (var_count) var_count[s] = a;
        gain_cnt[s] = b;
    }
}


// Gain recompute ("regain"): when a pick newly covers very many rows, subtracting them bit by bit costs one
// atomic per set bit (tens of millions for the first picks, which carry the common variants).  Recomputing
// every gain from the sample-major copy is one streaming pass (popcount of col & live) whatever the rows hold.
// Runs only when the selection kernel left st->regain set.
__global__ void __launch_bounds__(256) regain_kernel(SelParams p)
{
    if (p.st->regain == 0) return;
    const int s = blockIdx.x;
    if (s >= p.S) return;
    const uint4 *col = reinterpret_cast<const uint4 *>(p.cols + (long long)s * p.colPitchW);
    const uint4 *lv = reinterpret_cast<const uint4 *>(p.live);
    const long long n4 = p.colPitchW / 4;
    unsigned int alive = 0;
    unsigned long long lo = 0, hi = 0;
    for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
        const uint4 c = ld_stream_u128(col + i);
        const uint4 l = __ldcg(lv + i);
        uint32_t w[4] = {c.x & l.x, c.y & l.y, c.z & l.z, c.w & l.w};
        alive += __popc(w[0]) + __popc(w[1]) + __popc(w[2]) + __popc(w[3]);
        if (p.af) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                uint32_t x = w[u];
                while (x) {
                    const long long r = (i * 4 + u) * 32 + (__ffs(x) - 1);
                    x &= x - 1;
                    lo += __ldg(p.q_lo + r);
                    hi += __ldg(p.q_hi + r);
                }
            }
        }
    }
    __shared__ unsigned int s_alive[8];
    __shared__ unsigned long long s_lo[8], s_hi[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        alive += __shfl_xor_sync(0xffffffffu, alive, o);
        lo += __shfl_xor_sync(0xffffffffu, lo, o);
        hi += __shfl_xor_sync(0xffffffffu, hi, o);
    }
    if ((threadIdx.x & 31) == 0) { s_alive[threadIdx.x >> 5] = alive; s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int a = 0;
        unsigned long long l = 0, h = 0;
        for (int i = 0; i < 8; ++i) { a += s_alive[i]; l += s_lo[i]; h += s_hi[i]; }
        p.gain_cnt[s] = a;
        if (p.af) { p.gain_lo[s] = l; p.gain_hi[s] = h; }
    }
}

// K5 for a heavy pick in count mode ("incremental gain decrement", utmos/select.py:37-41 restricted to the rows the
// pick newly covered): gain_cnt[s] -= number of those rows that carry s.  The rows are read once from the
// variant-major matrix (N_t * pitch bytes instead of the whole sample-major copy that regain_kernel streams) and
// counted by positional popcount: a thread owns one word column (32 samples) and adds 16 rows at a time into
// bit-sliced counters (plane k holds bit k of the 32 counts) with a carry-save adder tree; at the end the 16 planes
// are turned into 32 integers by one 32x32 bit transpose in registers, the slices of a CTA are summed in shared
// memory and every sample gets one RED per CTA.  Runs only when the selection kernel left st->regain set.
constexpr int kDecBatchWords = 64;       // mask words whose rows are listed at a time (2,048 rows)
constexpr int kDecMaxWords = 1984;       // mask words per CTA: fewer than 65,536 rows, so 16 planes cannot overflow

__device__ __forceinline__ void csa(uint32_t &h, uint32_t &l, uint32_t a, uint32_t b, uint32_t c)
{
    const uint32_t u = a ^ b;
    h = (a & b) | (u & c);
    l = u ^ c;
}

__global__ void __launch_bounds__(256) cover_decrement_kernel(SelParams p, const uint32_t *__restrict__ newmask, int words_per_cta)
{
    if (p.st->regain == 0) return;
    __shared__ unsigned int s_rows[kDecBatchWords * 32];
    __shared__ unsigned int s_n;
    extern __shared__ unsigned int s_dec[];                      // [cols_here * 32] decrements found by this CTA
    const int tid = threadIdx.x;
    const int col0 = blockIdx.y * 256;
    const int cols_here = min(256, p.pitchW - col0);             // word columns of this CTA
    const int slices = 256 / cols_here;                          // row slices working side by side on those columns
    const int wc = tid % cols_here, sl = tid / cols_here;
    const bool active = sl < slices;
    for (int i = tid; i < cols_here * 32; i += 256) s_dec[i] = 0u;
    uint32_t pl[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) pl[k] = 0u;
    const long long mw0 = (long long)blockIdx.x * words_per_cta;
    const long long mw1 = min(mw0 + (long long)words_per_cta, p.colPitchW);
    for (long long mb = mw0; mb < mw1; mb += kDecBatchWords) {
        __syncthreads();                                         // the previous list has been consumed
        if (tid == 0) s_n = 0u;
        __syncthreads();
        if (tid < kDecBatchWords && mb + tid < mw1) {
            uint32_t x = __ldcg(newmask + mb + tid);
            if (x) {
                unsigned int at = atomicAdd(&s_n, (unsigned int)__popc(x));
                const unsigned int r0 = (unsigned int)((mb + tid) << 5);
                while (x) {
                    s_rows[at++] = r0 + (unsigned int)(__ffs(x) - 1);
                    x &= x - 1;
                }
            }
        }
        __syncthreads();
        const int n = (int)s_n;
        if (!active) continue;
        for (int i0 = sl * 16; i0 < n; i0 += slices * 16) {
            uint32_t x[16];
#pragma unroll
            for (int u = 0; u < 16; ++u)
                x[u] = i0 + u < n ? __ldg(p.rows + (long long)s_rows[i0 + u] * p.pitchW + col0 + wc) : 0u;
            uint32_t t2a, t2b, t4a, t4b, t8a, t8b, t16;
            csa(t2a, pl[0], pl[0], x[0], x[1]);
            csa(t2b, pl[0], pl[0], x[2], x[3]);
            csa(t4a, pl[1], pl[1], t2a, t2b);
            csa(t2a, pl[0], pl[0], x[4], x[5]);
            csa(t2b, pl[0], pl[0], x[6], x[7]);
            csa(t4b, pl[1], pl[1], t2a, t2b);
            csa(t8a, pl[2], pl[2], t4a, t4b);
            csa(t2a, pl[0], pl[0], x[8], x[9]);
            csa(t2b, pl[0], pl[0], x[10], x[11]);
            csa(t4a, pl[1], pl[1], t2a, t2b);
            csa(t2a, pl[0], pl[0], x[12], x[13]);
            csa(t2b, pl[0], pl[0], x[14], x[15]);
            csa(t4b, pl[1], pl[1], t2a, t2b);
            csa(t8b, pl[2], pl[2], t4a, t4b);
            csa(t16, pl[3], pl[3], t8a, t8b);
#pragma unroll
            for (int k = 4; k < 16; ++k) {                       // add the carry of weight 16 into the upper planes
                const uint32_t t = pl[k] & t16;
                pl[k] ^= t16;
                t16 = t;
            }
        }
    }
    if (active) {
        // 16 planes -> 32 counts: word b of the transposed 32x32 bit block has bit k = bit b of plane k
        uint32_t a[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) a[k] = k < 16 ? pl[k] : 0u;
        transpose32_regs(a);
#pragma unroll
        for (int b = 0; b < 32; ++b)
            if (a[b]) atomicAdd(&s_dec[wc * 32 + b], a[b]);
    }
    __syncthreads();
    for (int i = tid; i < cols_here * 32; i += 256) {
        const int s = col0 * 32 + i;
        const unsigned int d = s_dec[i];
        if (d && s < p.S) atomicSub(p.gain_cnt + s, d);
    }
}

// K3, AF limbs (sample-major source): gain_lo/hi[s] = sum of q_lo/hi[r] over the live rows r that carry s
// (the step-0 scores of utmos/select.py:37-40 with data = GT * AF, :317-320, as exact fixed-point integers).
// CTA = (tile of kAfRows rows) x (group of kAfSamples samples).  The limbs of the tile's rows are staged in shared
// memory ONCE (16 bytes per row) instead of being gathered from L2 once per set bit; a thread then takes whole
// samples: it reads the sample's 256 contiguous bytes of the tile with 128-bit loads and adds the staged limbs of its
// set bits into registers.  One pair of 64-bit atomics per (sample, tile).  Algorithmic bytes: V'*pitch + 16*V'.
constexpr int kAfRows = 2048;
constexpr int kAfWords = kAfRows / 32;
constexpr int kAfSamples = 512;

__global__ void __launch_bounds__(256) col_af_kernel(const uint32_t *__restrict__ cols, const uint32_t *__restrict__ live,
                                                     const unsigned long long *__restrict__ q_lo,
                                                     const unsigned long long *__restrict__ q_hi, long long colPitchW,
                                                     long long V, int S, unsigned long long *gain_lo,
                                                     unsigned long long *gain_hi)
{
    __shared__ ulonglong2 s_q[kAfRows];
    __shared__ uint32_t s_live[kAfWords];
    const long long w0 = (long long)blockIdx.x * kAfWords;
    const long long r0 = w0 * 32;
    for (int i = threadIdx.x; i < kAfRows; i += blockDim.x) {
        const long long r = r0 + i;
        s_q[i] = r < V ? make_ulonglong2(__ldg(q_lo + r), __ldg(q_hi + r)) : make_ulonglong2(0ull, 0ull);
    }
    if (threadIdx.x < kAfWords) s_live[threadIdx.x] = w0 + threadIdx.x < colPitchW ? __ldg(live + w0 + threadIdx.x) : 0u;
    __syncthreads();
    const int s_end = min(S, (int)(blockIdx.y + 1) * kAfSamples);
    for (int s = (int)blockIdx.y * kAfSamples + threadIdx.x; s < s_end; s += blockDim.x) {
        const uint4 *col = reinterpret_cast<const uint4 *>(cols + (long long)s * colPitchW + w0);
        unsigned long long lo = 0ull, hi = 0ull;
#pragma unroll 4
        for (int j = 0; j < kAfWords / 4; ++j) {
            if (w0 + 4 * j >= colPitchW) break;                 // colPitchW is a multiple of 8 words
            const uint4 c = ld_stream_u128(col + j);
            const uint32_t w[4] = {c.x & s_live[4 * j], c.y & s_live[4 * j + 1], c.z & s_live[4 * j + 2], c.w & s_live[4 * j + 3]};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t x = w[k];
                while (x) {
                    const ulonglong2 q = s_q[(4 * j + k) * 32 + (__ffs(x) - 1)];
                    x &= x - 1;
                    lo += q.x;
                    hi += q.y;
                }
            }
        }
        if (lo | hi) {
            atomicAdd(gain_lo + s, lo);
            atomicAdd(gain_hi + s, hi);
        }
    }
}

// K3 (variant-major source): one warp per row, one reduction per set bit.
//   WHAT bit 0: var_count over all rows, bit 1: gain_cnt over live rows, bit 2: AF limbs over live rows
__global__ void __launch_bounds__(256) row_gain_kernel(SelParams p, unsigned int *var_count, int what)
{
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = warp0; r < p.V; r += nwarps) {
        const bool alive = (p.live[r >> 5] >> (r & 31)) & 1u;
        if (!(what & 1) && !alive) continue;
        const uint32_t *row = p.rows + r * p.pitchW;
        unsigned long long ql = 0, qh = 0;
        if ((what & 4) && alive) { ql = p.q_lo[r]; qh = p.q_hi[r]; }
        for (int k = lane; k < p.nW; k += 32) {
            uint32_t x = row[k];
            while (x) {
                const int s = k * 32 + (__ffs(x) - 1);
                x &= x - 1;
                if (what & 1) atomicAdd(var_count + s, 1u);
                if (alive) {
                    if (what & 2) atomicAdd(p.gain_cnt + s, 1u);
                    if (what & 4) { atomicAdd(p.gain_lo + s, ql); atomicAdd(p.gain_hi + s, qh); }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K4 building blocks
// ------------------------------------------------------------------------------------------------
struct Best {
    double score;
    int idx;
    unsigned int cnt;
};

// post-mask, post-weight score of sample s (utmos/select.py:43-47) and its current new_count
__device__ __forceinline__ void sample_score(const SelParams &p, int s, double *score, unsigned int *cnt)
{
    const unsigned int c = __ldcg(p.gain_cnt + s);
    double g = 0.0;
    if (__ldcg(p.mask + s) == 1) {
        g = p.af ? fixed_to_double(__ldcg(p.gain_lo + s), __ldcg(p.gain_hi + s), p.L, p.scale) : (double)c;
        if (p.weights) g *= __ldg(p.weights + s);
    }
    *score = g;
    *cnt = c;
}

__device__ __forceinline__ Best best_of(Best a, Best b) { return arg_better(b.score, b.idx, a.score, a.idx) ? b : a; }

__device__ __forceinline__ Best warp_best(Best v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Best t;
        t.score = __shfl_xor_sync(0xffffffffu, v.score, o);
        t.idx = __shfl_xor_sync(0xffffffffu, v.idx, o);
        t.cnt = __shfl_xor_sync(0xffffffffu, v.cnt, o);
        v = best_of(v, t);
    }
    return v;
}

// CTA-wide reduction; result valid in every thread.  s_red must hold 32 entries.
__device__ __forceinline__ Best block_best(Best v, Best *s_red)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    v = warp_best(v);
    __syncthreads();                       // s_red reuse
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    Best t = lane < nwarp ? s_red[lane] : Best{-1.0e308, 0x7fffffff, 0u};
    // "empty" entries lose against anything real: score of -1e308 never beats a finite score, and on
    // equality the index 0x7fffffff loses
    t = warp_best(t);
    return t;
}

__device__ __forceinline__ Best scan_best(const SelParams &p, int begin, int end)
{
    Best b{-1.0e308, 0x7fffffff, 0u};
    for (int s = begin + (int)threadIdx.x; s < end; s += (int)blockDim.x) {
        double sc;
        unsigned int c;
        sample_score(p, s, &sc, &c);
        if (arg_better(sc, s, b.score, b.idx)) { b.score = sc; b.idx = s; b.cnt = c; }
    }
    return b;
}

// ------------------------------------------------------------------------------------------------
// K5 building blocks
// ------------------------------------------------------------------------------------------------
// subtract row r from every sample that carries it (whole warp cooperates on one row)
__device__ __forceinline__ void retire_row(const SelParams &p, long long r, int lane)
{
    const uint32_t *row = p.rows + r * p.pitchW;
    unsigned long long nl = 0, nh = 0;
    if (p.af) { nl = 0ull - p.q_lo[r]; nh = 0ull - p.q_hi[r]; }
    for (int k = lane; k < p.nW; k += 32) {
        uint32_t x = __ldg(row + k);
        while (x) {
            const int s = k * 32 + (__ffs(x) - 1);
            x &= x - 1;
            atomicAdd(p.gain_cnt + s, 0xffffffffu);
            if (p.af) { atomicAdd(p.gain_lo + s, nl); atomicAdd(p.gain_hi + s, nh); }
        }
    }
}

// one warp handles 32 consecutive words (1024 rows) of the live mask for winner b
__device__ __forceinline__ void cover_chunk(const SelParams &p, int b, long long chunk, int lane, bool retire = true)
{
    const long long w = chunk * 32 + lane;
    uint32_t lv = 0, nw = 0;
    if (w < p.colPitchW) lv = __ldcg(p.live + w);
    if (p.cols) {
        if (w < p.colPitchW) nw = lv & __ldg(p.cols + (long long)b * p.colPitchW + w);
    } else if (lv) {
        // no sample-major copy: probe the winner's bit of every live row (one 32 B sector per row)
        const int bw = b >> 5, bb = b & 31;
#pragma unroll 8
        for (int j = 0; j < 32; ++j) {
            if ((lv >> j) & 1u) {
                const uint32_t x = __ldg(p.rows + (w * 32 + j) * p.pitchW + bw);
                nw |= ((x >> bb) & 1u) << j;
            }
        }
    }
    if (nw) __stcg(p.live + w, lv ^ nw);
    if (!retire) return;                              // gains will be recomputed by regain_kernel
    unsigned int pending = __ballot_sync(0xffffffffu, nw != 0);
    while (pending) {
        const int src = __ffs(pending) - 1;
        pending &= pending - 1;
        uint32_t m = __shfl_sync(0xffffffffu, nw, src);
        const long long rbase = (chunk * 32 + src) * 32;
        while (m) {
            const int j = __ffs(m) - 1;
            m &= m - 1;
            retire_row(p, rbase + j, lane);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// step-kernel flavour (UTMOS_F_STEP_KERNELS): two launches per step, replayed from a CUDA graph
// ------------------------------------------------------------------------------------------------
// UTMOS_F_REF_TIES (--af only): the reference adds the float64 AF of the uncovered rows that carry a sample one after the
// other in row order (utmos/select.py:37-40), so two samples whose exact sums tie -- or nearly tie -- are ordered by the
// rounding noise of those sequential sums.  The exact fixed-point scores decide everything else; for the candidates within
// 2^-30 relative of the best exact score (the reference's accumulated error is below V * 2^-53 <= 2^-32) one warp per
// candidate replays the reference's sum: the live rows of its sample-major row in ascending order, one float64 add each
// (the lanes locate the set bits, the adds stay strictly sequential), times the weight (:47); the first maximum wins (:48).
__device__ __forceinline__ double replay_reference_sum(const SelParams &p, int t, int lane)
{
    // a lane takes four consecutive words (128 rows) per turn with one 128-bit load of the column and of the live mask
    // (colPitchW is a multiple of 8 words and the words past V are zero in both); lanes are served in ascending order,
    // a lane adds its rows in ascending order: the reference's row order
    const uint4 *col = reinterpret_cast<const uint4 *>(p.cols + (long long)t * p.colPitchW);
    const uint4 *lv = reinterpret_cast<const uint4 *>(p.live);
    const long long n4 = p.colPitchW >> 2;
    double acc = 0.0;
    for (long long q0 = 0; q0 < n4; q0 += 32) {
        const long long q = q0 + lane;
        uint4 x = make_uint4(0u, 0u, 0u, 0u);
        if (q < n4) {
            const uint4 c = __ldg(col + q);
            const uint4 l = __ldcg(lv + q);
            x = make_uint4(c.x & l.x, c.y & l.y, c.z & l.z, c.w & l.w);
        }
        unsigned int m = __ballot_sync(0xffffffffu, (x.x | x.y | x.z | x.w) != 0u);
        while (m) {
            const int ln = __ffs(m) - 1;
            m &= m - 1;
            if (lane == ln) {
                double a = acc;
                const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t y = w[k];
                    while (y) {
                        const long long r = (((q << 2) + k) << 5) + (__ffs(y) - 1);
                        y &= y - 1;
                        double v = __ldg(p.af_vals + r);
                        if (p.af_f32) v = (double)(float)v;      // hdf5 flavour: float32 GT*AF rows (utmos/select.py:218-223)
                        a += v;
                    }
                }
                acc = a;
            }
            acc = __shfl_sync(0xffffffffu, acc, ln);
        }
    }
    return acc;
}

__global__ void __launch_bounds__(1024) argmax_step_kernel(SelParams p)
{
    __shared__ Best s_red[32];
    __shared__ int s_cand[1024];
    __shared__ int s_ncand;
    SelState *st = p.st;
    const bool idle = st->stop != 0 || st->step >= st->limit;
    if (threadIdx.x == 0) s_ncand = 0;
    __syncthreads();
    if (idle) {
        if (threadIdx.x == 0) st->winner = -1;
        return;
    }
    Best b = block_best(scan_best(p, 0, p.S), s_red);
    if (p.ref_ties && p.af && p.cols && p.af_vals && b.score > 0.0) {
        const double thr = b.score * (1.0 - 9.313225746154785e-10);        // 2^-30
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        // how many candidates are there at all?  (one: the exact winner stands)
        int mine = 0;
        for (int s = (int)threadIdx.x; s < p.S; s += 1024) {
            double sc;
            unsigned int c;
            sample_score(p, s, &sc, &c);
            mine += sc >= thr ? 1 : 0;
        }
        const int holders = __syncthreads_count(mine > 0);                  // threads that hold a candidate
        const bool several = __syncthreads_or(mine > 1) || holders > 1;
        if (several) {
            if (threadIdx.x == 0 && p.dbg) p.dbg[12] += 1;                  // steps decided by the replay
            Best rb{-1.0e308, 0x7fffffff, 0u};
            for (int base = 0; base < p.S; base += 1024) {                 // at most 1,024 candidates per round
                const int s = base + (int)threadIdx.x;
                if (s < p.S) {
                    double sc;
                    unsigned int c;
                    sample_score(p, s, &sc, &c);
                    if (sc >= thr) s_cand[atomicAdd(&s_ncand, 1)] = s;
                }
                __syncthreads();
                const int n = s_ncand;
                for (int i = warp; i < n; i += 32) {
                    const int t = s_cand[i];
                    double f = replay_reference_sum(p, t, lane);
                    if (p.weights) f *= __ldg(p.weights + t);
                    if (lane == 0 && arg_better(f, t, rb.score, rb.idx)) { rb.score = f; rb.idx = t; rb.cnt = __ldcg(p.gain_cnt + t); }
                }
                __syncthreads();
                if (threadIdx.x == 0) s_ncand = 0;
                __syncthreads();
            }
            if (lane != 0) rb = Best{-1.0e308, 0x7fffffff, 0u};
            b = block_best(rb, s_red);
        }
    }
    if (threadIdx.x == 0) {
        if (p.S == 0 || b.score == 0.0) {                 // utmos/select.py:51-52
            st->stop = UTMOS_STOP_ZERO;
            st->winner = -1;
        } else {
            const long long i = st->step;
            p.out_idx[i] = b.idx;
            p.out_new[i] = b.cnt;
            p.out_score[i] = b.score;
            if (p.dbg_time) p.out_time[i] = global_timer_ns();
            st->step = i + 1;
            st->tot += b.cnt;
            p.mask[b.idx] = 0;                            // utmos/select.py:100
            st->winner = b.idx;
            if (st->tot >= p.V) {                         // utmos/select.py:110-112
                st->stop = UTMOS_STOP_ALL;
                st->winner = -1;
            }
        }
    }
}

__global__ void __launch_bounds__(256) cover_step_kernel(SelParams p)
{
    const int b = p.st->winner;
    if (b < 0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) p.st->regain = 0;
        return;
    }
    // a pick that newly covers >= regain_rows rows (0 = never): only the live bits are cleared here, the gains are
    // recomputed by the regain_kernel that follows (one streaming pass instead of one atomic per set bit of those rows)
    const bool heavy = p.regain_rows && p.cols && p.out_new[p.st->step - 1] >= (long long)p.regain_rows;
    if (blockIdx.x == 0 && threadIdx.x == 0) p.st->regain = heavy ? 1u : 0u;
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long nchunks = (p.colPitchW + 31) / 32;
    for (long long c = warp0; c < nchunks; c += nwarps) cover_chunk(p, b, c, lane, !heavy);
}

// ------------------------------------------------------------------------------------------------
// persistent flavour: one cooperative launch runs the whole selection; CTAs meet at a grid barrier
// twice per step (after the partial argmax, after the cover phase).
// ------------------------------------------------------------------------------------------------
constexpr long long kSpinLimit = 1ll << 27;     // watchdog: ~seconds, far above any legitimate wait

__device__ __forceinline__ bool grid_barrier(unsigned int *counter, unsigned int *epoch, unsigned int nblocks,
                                             unsigned int *abort_flag, int *s_flag)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        *epoch += 1;
        const unsigned int target = *epoch * nblocks;
        __threadfence();
        atomicAdd(counter, 1u);
        long long spins = 0;
        int bad = 0;
        while ((int)(ld_acquire_u32(counter) - target) < 0) {
            if (++spins > kSpinLimit) { atomicExch(abort_flag, 1u); bad = 1; break; }
            if ((spins & 0xfff) == 0 && ld_acquire_u32(abort_flag)) { bad = 1; break; }
        }
        if (!bad && ld_acquire_u32(abort_flag)) bad = 1;
        __threadfence();
        *s_flag = bad;
    }
    __syncthreads();
    return *s_flag == 0;
}

__global__ void __launch_bounds__(1024, 1) select_persistent_kernel(SelParams p, unsigned int *bar_counter,
                                                                    ArgPartial *partials)
{
    __shared__ Best s_red[32];
    __shared__ int s_flag;
    __shared__ unsigned int s_epoch;
    SelState *st = p.st;
    if (threadIdx.x == 0) s_epoch = 0;
    const int lane = threadIdx.x & 31;
    const unsigned int nblocks = gridDim.x;
    // every CTA keeps the same private copy of the loop state (identical decisions everywhere)
    long long step = st->step, tot = st->tot;
    const long long limit = st->limit;
    int stop = st->stop;
    int regain = 0;
    int want_tail = 0;
    __shared__ unsigned long long s_sumw[32];
    const int per = (p.S + (int)nblocks - 1) / (int)nblocks;
    const int my_begin = min(p.S, (int)blockIdx.x * per), my_end = min(p.S, my_begin + per);
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)nblocks * blockDim.x) >> 5;
    const long long nchunks = (p.colPitchW + 31) / 32;
    __syncthreads();

    while (stop == 0 && step < limit) {
        // ---- phase A: partial argmax over this CTA's slice of the samples (+ sum of its gains)
        {
            unsigned long long acc = 0;
            if (p.tail_budget)
                for (int s = my_begin + (int)threadIdx.x; s < my_end; s += (int)blockDim.x) acc += __ldcg(p.gain_cnt + s);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) s_sumw[threadIdx.x >> 5] = acc;
        }
        Best b = block_best(scan_best(p, my_begin, my_end), s_red);
        if (threadIdx.x == 0) {
            ArgPartial a;
            a.score = b.score; a.idx = b.idx; a.cnt = b.cnt;
            a.sum = 0;
            for (int i = 0; i < (int)(blockDim.x >> 5); ++i) a.sum += s_sumw[i];
            partials[blockIdx.x] = a;
        }
        if (!grid_barrier(bar_counter, &s_epoch, nblocks, &st->abort_flag, &s_flag)) return;
        Best t{-1.0e308, 0x7fffffff, 0u};
        unsigned long long live_now = 0;
        for (unsigned int i = threadIdx.x; i < nblocks; i += blockDim.x) {
            Best o;
            o.score = __ldcg(&partials[i].score);
            o.idx = __ldcg(&partials[i].idx);
            o.cnt = __ldcg(&partials[i].cnt);
            live_now += __ldcg(&partials[i].sum);
            t = best_of(t, o);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) live_now += __shfl_xor_sync(0xffffffffu, live_now, o);
        __syncthreads();
        if (lane == 0) s_sumw[threadIdx.x >> 5] = live_now;
        b = block_best(t, s_red);
        live_now = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) live_now += s_sumw[i];
        if (p.S == 0 || b.score == 0.0) {                  // utmos/select.py:51-52
            stop = UTMOS_STOP_ZERO;
            break;
        }
        if (p.tail_budget && live_now <= p.tail_budget && b.cnt < p.tail_rows) {
            want_tail = 1;                                 // sparse and small: the single-CTA tail kernel is faster
            break;
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            p.out_idx[step] = b.idx;
            p.out_new[step] = b.cnt;
            p.out_score[step] = b.score;
            if (p.dbg_time) p.out_time[step] = global_timer_ns();
            p.mask[b.idx] = 0;                             // utmos/select.py:100
        }
        step += 1;
        tot += b.cnt;
        if (tot >= p.V) {                                  // utmos/select.py:110-112
            stop = UTMOS_STOP_ALL;
            break;
        }
        // ---- phase B: clear the newly covered rows, subtract them from every gain (or leave that to regain)
        regain = p.cols && p.regain_rows && b.cnt >= p.regain_rows;
        for (long long c = warp0; c < nchunks; c += nwarps) cover_chunk(p, b.idx, c, lane, !regain);
        if (!grid_barrier(bar_counter, &s_epoch, nblocks, &st->abort_flag, &s_flag)) return;
        if (regain) break;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->step = step;
        st->tot = tot;
        st->stop = stop;
        st->winner = -1;
        st->regain = regain;
        st->want_tail = want_tail;
    }
}


// ------------------------------------------------------------------------------------------------
// cluster flavour (default when the state fits): ONE thread-block cluster of up to 16 CTAs runs the whole
// selection.  The step loop is latency bound (two dependent synchronisations per greedy step), so the
// grid-wide barrier through L2 is replaced by the hardware cluster barrier, and everything the loop
// touches every step lives in distributed shared memory:
//     gains, mask and weights of sample s  -> CTA s / per          (argmax never leaves the SM)
//     live-mask words                       -> CTA w / liveW        (probe = smem AND global column word)
// Decrements are fire-and-forget atomics on the owner CTA's shared memory (DSMEM).  Global memory
// traffic per step is the winner's column slice (one pass) plus the newly covered rows (each once).
// While a step runs, every CTA prefetches its slice of the runner-up's column into L2.
// ------------------------------------------------------------------------------------------------
namespace cg = cooperative_groups;

struct ClusterCfg {
    int per;            // samples per CTA (multiple of 32)
    int liveW;          // live-mask words per CTA (multiple of 32)
    int off_lo, off_hi, off_w, off_mask, off_live, off_part;   // byte offsets into dynamic smem (cnt at 0)
    int lanes_per_row;  // power of two: lanes that share one row when retiring
    int gains_l2;       // 1: gains stay in global memory (L2 atomics); 0: distributed shared memory (DSMEM atomics)
    uint32_t *newmask;  // [colPitchW] or null: the pick that ends the launch (st->regain) leaves the rows it newly covered
                        // here, for cover_decrement_kernel
};


__device__ __forceinline__ void retire_row_dsmem(const SelParams &p, const ClusterCfg &cfg, cg::cluster_group &cluster,
                                                 unsigned int *s_cnt, unsigned long long *s_lo,
                                                 unsigned long long *s_hi, long long r, int lane)
{
    const uint32_t *row = p.rows + r * p.pitchW;
    const int wpc = cfg.per >> 5;                 // words per owner CTA
    unsigned long long nl = 0, nh = 0;
    if (p.af) { nl = 0ull - p.q_lo[r]; nh = 0ull - p.q_hi[r]; }
    for (int k0 = 0; k0 < p.nW; k0 += 128) {
        uint32_t x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = k0 + u * 32 + lane;
            x[u] = k < p.nW ? __ldg(row + k) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            uint32_t w = x[u];
            if (!w) continue;
            const int k = k0 + u * 32 + lane;
            const int owner = k / wpc;
            const int base = (k - owner * wpc) << 5;
            unsigned int *rc;
            unsigned long long *rl = nullptr, *rh = nullptr;
            if (cfg.gains_l2) {
                rc = p.gain_cnt + (k << 5);
                if (p.af) { rl = p.gain_lo + (k << 5); rh = p.gain_hi + (k << 5); }
            } else {
                rc = cluster.map_shared_rank(s_cnt, owner) + base;
                if (p.af) { rl = cluster.map_shared_rank(s_lo, owner) + base; rh = cluster.map_shared_rank(s_hi, owner) + base; }
            }
            while (w) {
                const int j = __ffs(w) - 1;
                w &= w - 1;
                atomicAdd(rc + j, 0xffffffffu);
                if (p.af) { atomicAdd(rl + j, nl); atomicAdd(rh + j, nh); }
            }
        }
    }
}

// NEWMASK (opt-in): the pick that ends the launch leaves the bitmask of its newly covered rows in cfg.newmask.
template <bool NEWMASK>
__global__ void __launch_bounds__(1024, 1) select_cluster_kernel(SelParams p, ClusterCfg cfg)
{
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ Best s_red[32];
    unsigned int *s_cnt = reinterpret_cast<unsigned int *>(smem);
    unsigned long long *s_lo = reinterpret_cast<unsigned long long *>(smem + cfg.off_lo);
    unsigned long long *s_hi = reinterpret_cast<unsigned long long *>(smem + cfg.off_hi);
    double *s_w = reinterpret_cast<double *>(smem + cfg.off_w);
    uint8_t *s_mask = smem + cfg.off_mask;
    uint32_t *s_live = reinterpret_cast<uint32_t *>(smem + cfg.off_live);
    ArgPartial *s_part = reinterpret_cast<ArgPartial *>(smem + cfg.off_part);

    const int rank = (int)cluster.block_rank();
    const int CL = (int)cluster.num_blocks();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    const int s0 = rank * cfg.per;
    const int n_mine = max(0, min(cfg.per, p.S - s0));
    const long long w0 = (long long)rank * cfg.liveW;
    const bool has_w = p.weights != nullptr;
    SelState *st = p.st;

    // ---- load this CTA's slice of the state into shared memory
    for (int i = tid; i < cfg.per; i += blockDim.x) {
        const bool ok = i < n_mine;
        s_mask[i] = ok ? p.mask[s0 + i] : (uint8_t)2;
        if (!cfg.gains_l2) {
            s_cnt[i] = ok ? p.gain_cnt[s0 + i] : 0u;
            if (p.af) { s_lo[i] = ok ? p.gain_lo[s0 + i] : 0ull; s_hi[i] = ok ? p.gain_hi[s0 + i] : 0ull; }
        }
        if (has_w) s_w[i] = ok ? p.weights[s0 + i] : 0.0;
    }
    for (int i = tid; i < cfg.liveW; i += blockDim.x) s_live[i] = w0 + i < p.colPitchW ? p.live[w0 + i] : 0u;
    long long step = st->step, tot = st->tot;
    const long long limit = st->limit;
    int stop = st->stop;
    int regain = 0, want_tail = 0;
    const int nchunks = cfg.liveW >> 5;
    __shared__ unsigned long long s_sumw[32];
    __shared__ unsigned int s_wq[32][64];               // per-warp queue of rows to retire
    cluster.sync();

    long long t_a = 0, t_s1 = 0, t_b = 0, t_s2 = 0, t_mark = clock64();
#define UT_TICK(acc) do { const long long now__ = clock64(); acc += now__ - t_mark; t_mark = now__; } while (0)
    while (stop == 0 && step < limit) {
        // ---- phase A: argmax over the samples this CTA owns (shared memory only)
        Best b{-1.0e308, 0x7fffffff, 0u};
        unsigned long long acc = 0;
        for (int i = tid; i < n_mine; i += blockDim.x) {
            const unsigned int c = cfg.gains_l2 ? __ldcg(p.gain_cnt + s0 + i) : s_cnt[i];
            acc += c;
            double g = 0.0;
            if (s_mask[i] == 1) {
                if (p.af) {
                    const unsigned long long lo = cfg.gains_l2 ? __ldcg(p.gain_lo + s0 + i) : s_lo[i];
                    const unsigned long long hi = cfg.gains_l2 ? __ldcg(p.gain_hi + s0 + i) : s_hi[i];
                    g = fixed_to_double(lo, hi, p.L, p.scale);
                } else {
                    g = (double)c;
                }
                if (has_w) g *= s_w[i];
            }
            if (arg_better(g, s0 + i, b.score, b.idx)) { b.score = g; b.idx = s0 + i; b.cnt = c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) s_sumw[warp] = acc;
        b = block_best(b, s_red);
        if (tid < CL) {                                   // publish the partial in every CTA of the cluster
            ArgPartial *remote = cluster.map_shared_rank(s_part, tid);
            ArgPartial a;
            a.score = b.score; a.idx = b.idx; a.cnt = b.cnt;
            a.sum = 0;
            for (int i = 0; i < nwarp; ++i) a.sum += s_sumw[i];
            remote[rank] = a;
        }
        UT_TICK(t_a);
        cluster.sync();
        UT_TICK(t_s1);
        // every warp reduces the CL partials redundantly (no further CTA-wide sync needed)
        Best t{-1.0e308, 0x7fffffff, 0u};
        unsigned long long live_now = 0;
        if (lane < CL) { t.score = s_part[lane].score; t.idx = s_part[lane].idx; t.cnt = s_part[lane].cnt; live_now = s_part[lane].sum; }
        const Best mine = t;
        b = warp_best(t);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) live_now += __shfl_xor_sync(0xffffffffu, live_now, o);
        if (p.S == 0 || b.score == 0.0) {                 // utmos/select.py:51-52
            stop = UTMOS_STOP_ZERO;
            break;
        }
        if (p.tail_budget && live_now <= p.tail_budget && b.cnt < p.tail_rows) {
            want_tail = 1;                                // sparse and small: the single-CTA tail kernel is faster
            break;
        }
        // runner-up among the other CTAs' local winners: a likely next pick -> prefetch its column slice
        Best t2 = (lane < CL && mine.idx != b.idx) ? mine : Best{-1.0e308, 0x7fffffff, 0u};
        t2 = warp_best(t2);
        const int owner = b.idx / cfg.per;
        if (tid == 0) {
            if (rank == 0) {
                p.out_idx[step] = b.idx;
                p.out_new[step] = b.cnt;
                p.out_score[step] = b.score;
                if (p.dbg_time) p.out_time[step] = global_timer_ns();
            }
            if (rank == owner) s_mask[b.idx - s0] = 0;    // utmos/select.py:100
        }
        step += 1;
        tot += b.cnt;
        if (tot >= p.V) {                                 // utmos/select.py:110-112
            stop = UTMOS_STOP_ALL;
            break;
        }
        // ---- phase B: clear newly covered rows from this CTA's live words, retire them everywhere
        regain = p.cols && p.regain_rows && b.cnt >= p.regain_rows;   // too many rows: recompute instead
        if (!regain && p.cols && t2.idx != 0x7fffffff && t2.score > 0.0) {
            const uint32_t *c2 = p.cols + (long long)t2.idx * p.colPitchW + w0;
            for (int i = tid * 32; i < cfg.liveW && w0 + i < p.colPitchW; i += blockDim.x * 32) prefetch_l2(c2 + i);
        }
        for (int c = warp; c < nchunks; c += nwarp) {
            const int wl = c * 32 + lane;
            const long long w = w0 + wl;
            const uint32_t lv = s_live[wl];
            uint32_t nw = 0;
            if (p.cols) {
                if (w < p.colPitchW) nw = lv & __ldg(p.cols + (long long)b.idx * p.colPitchW + w);
            } else if (lv) {
                const int bw = b.idx >> 5, bb = b.idx & 31;
#pragma unroll 8
                for (int j = 0; j < 32; ++j) {
                    if ((lv >> j) & 1u) {
                        const uint32_t x = __ldg(p.rows + (w * 32 + j) * p.pitchW + bw);
                        nw |= ((x >> bb) & 1u) << j;
                    }
                }
            }
            if (nw) s_live[wl] = lv ^ nw;
            if (regain) {                                 // warp-uniform
                if (NEWMASK && w < p.colPitchW) cfg.newmask[w] = nw;
                continue;
            }
            if (nw) {
                uint32_t m = nw;                          // warm L2 with the rows this lane found
                while (m) {
                    const int j = __ffs(m) - 1;
                    m &= m - 1;
                    const char *rp = reinterpret_cast<const char *>(p.rows + (w * 32 + j) * p.pitchW);
                    const int lines = min(8, (p.pitchW * 4 + 127) >> 7);
                    for (int l = 0; l < lines; ++l) prefetch_l2(rp + l * 128);
                }
            }
            // queue this warp's newly covered rows, then retire them G lanes per row with every 128-bit load of
            // up to 32/G rows in flight before the first use
            int mine_n = __popc(nw), incl = mine_n;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int tt = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += tt;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            if (total == 0) continue;
            if (total <= 64) {
                int pos = incl - mine_n;
                uint32_t m = nw;
                while (m) {
                    const int j = __ffs(m) - 1;
                    m &= m - 1;
                    s_wq[warp][pos++] = (unsigned int)((w << 5) + j);
                }
                __syncwarp();
                const int G = cfg.lanes_per_row, sub = lane & (G - 1), slot = lane / G, rpw = 32 / G;
                const int n4 = p.pitchW >> 2, wpc = cfg.per >> 5;
                for (int q0 = 0; q0 < total; q0 += rpw) {
                    const int q = q0 + slot;
                    const bool have = q < total;
                    const long long r = have ? (long long)s_wq[warp][q] : 0ll;
                    const uint4 *row4 = reinterpret_cast<const uint4 *>(p.rows + r * p.pitchW);
                    unsigned long long nl = 0, nh = 0;
                    if (p.af && have) { nl = 0ull - p.q_lo[r]; nh = 0ull - p.q_hi[r]; }
                    for (int t0 = 0; t0 * G < n4; t0 += 4) {
                        uint4 x[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int j = (t0 + u) * G + sub;
                            x[u] = (have && j < n4) ? __ldg(row4 + j) : make_uint4(0u, 0u, 0u, 0u);
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int j = (t0 + u) * G + sub;
                            const uint32_t w4[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                uint32_t ww = w4[e];
                                if (!ww) continue;
                                const int k = j * 4 + e;
                                unsigned int *rc;
                                unsigned long long *rl = nullptr, *rh = nullptr;
                                if (cfg.gains_l2) {
                                    rc = p.gain_cnt + (k << 5);
                                    if (p.af) { rl = p.gain_lo + (k << 5); rh = p.gain_hi + (k << 5); }
                                } else {
                                    const int owner = k / wpc;
                                    const int sb = (k - owner * wpc) << 5;
                                    rc = cluster.map_shared_rank(s_cnt, owner) + sb;
                                    if (p.af) { rl = cluster.map_shared_rank(s_lo, owner) + sb; rh = cluster.map_shared_rank(s_hi, owner) + sb; }
                                }
                                while (ww) {
                                    const int jj = __ffs(ww) - 1;
                                    ww &= ww - 1;
                                    atomicAdd(rc + jj, 0xffffffffu);
                                    if (p.af) { atomicAdd(rl + jj, nl); atomicAdd(rh + jj, nh); }
                                }
                            }
                        }
                    }
                }
                __syncwarp();
            } else {
                unsigned int pending = __ballot_sync(0xffffffffu, nw != 0);
                while (pending) {
                    const int src = __ffs(pending) - 1;
                    pending &= pending - 1;
                    uint32_t m = __shfl_sync(0xffffffffu, nw, src);
                    const long long rbase = (w0 + c * 32 + src) * 32;
                    while (m) {
                        const int j = __ffs(m) - 1;
                        m &= m - 1;
                        retire_row_dsmem(p, cfg, cluster, s_cnt, s_lo, s_hi, rbase + j, lane);
                    }
                }
            }
        }
        UT_TICK(t_b);
        cluster.sync();
        UT_TICK(t_s2);
        if (regain) break;
    }
#undef UT_TICK
    if (rank == 0 && tid == 0 && p.dbg) {
        p.dbg[0] += t_a; p.dbg[1] += t_s1; p.dbg[2] += t_b; p.dbg[3] += t_s2; p.dbg[4] += 1;
    }

    // ---- write the state back so the selection can be resumed / inspected
    cluster.sync();
    for (int i = tid; i < n_mine; i += blockDim.x) {
        p.mask[s0 + i] = s_mask[i];
        if (!cfg.gains_l2) {
            p.gain_cnt[s0 + i] = s_cnt[i];
            if (p.af) { p.gain_lo[s0 + i] = s_lo[i]; p.gain_hi[s0 + i] = s_hi[i]; }
        }
    }
    for (int i = tid; i < cfg.liveW; i += blockDim.x)
        if (w0 + i < p.colPitchW) p.live[w0 + i] = s_live[i];
    if (rank == 0 && tid == 0) {
        st->step = step;
        st->tot = tot;
        st->stop = stop;
        st->winner = -1;
        st->regain = regain;
        st->want_tail = want_tail;
    }
}



// ------------------------------------------------------------------------------------------------
// multi-GPU flavour: one persistent cooperative kernel per GPU.  Rows are sharded over the ranks, the
// gains are replicated.  Every step each rank (A) finds the same winner from its replica, (B) retires
// the rows of ITS shard the winner newly covers into a local delta vector, (C) stores that delta
// straight into every peer's inbox over NVLink (P2P stores on IPC-mapped pointers) and publishes a
// sequence number, (D) waits for the peers' sequence numbers and adds all deltas to its replica.
// Integer (count) and fixed-point (AF limb) deltas make the result independent of the rank count.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

template <bool CLUSTER>
__device__ __forceinline__ bool mgpu_barrier(unsigned int *counter, unsigned int *epoch, unsigned int nblocks,
                                             unsigned int *abort_flag, int *s_flag)
{
    if (CLUSTER) {
        cg::this_cluster().sync();
        return true;
    }
    return grid_barrier(counter, epoch, nblocks, abort_flag, s_flag);
}

// CLUSTER: the CTAs of this GPU form ONE thread-block cluster and meet at the hardware cluster barrier (about 1 us)
// instead of a grid-wide barrier through L2 (about 5 us with the spin on one address); the loop has 4-5 barriers per step.
template <bool CLUSTER>
__global__ void __launch_bounds__(1024, 1) select_mgpu_kernel(SelParams p, MgpuParams m, unsigned int *bar_counter,
                                                              ArgPartial *partials)
{
    __shared__ Best s_red[32];
    __shared__ int s_flag;
    __shared__ unsigned int s_epoch;
    SelState *st = p.st;
    if (threadIdx.x == 0) s_epoch = 0;
    const int lane = threadIdx.x & 31;
    const unsigned int nblocks = gridDim.x;
    long long step = st->step, tot = st->tot;
    const long long limit = st->limit;
    int stop = st->stop;
    unsigned long long seq = m.seq0;
    const int per = (p.S + (int)nblocks - 1) / (int)nblocks;
    const int my_begin = min(p.S, (int)blockIdx.x * per), my_end = min(p.S, my_begin + per);
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nthreads = (long long)nblocks * blockDim.x;
    const long long warp0 = gtid >> 5, nwarps = nthreads >> 5;
    const long long nchunks = (p.colPitchW + 31) / 32;
    // the cover phase subtracts into the delta vectors instead of the gains
    SelParams pd = p;
    pd.gain_cnt = m.delta_cnt;
    pd.gain_lo = m.delta_lo;
    pd.gain_hi = m.delta_hi;
    __syncthreads();

    int want_tail = 0;
    __shared__ unsigned long long s_sumw[32];
    while (stop == 0 && step < limit) {
        // ---- A: replicated argmax (identical on every rank) + sum of the gains (= live set bits, all ranks)
        {
            unsigned long long acc = 0;
            if (p.tail_budget)
                for (int s = my_begin + (int)threadIdx.x; s < my_end; s += (int)blockDim.x) acc += __ldcg(p.gain_cnt + s);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) s_sumw[threadIdx.x >> 5] = acc;
        }
        Best b = block_best(scan_best(p, my_begin, my_end), s_red);
        if (threadIdx.x == 0) {
            ArgPartial a;
            a.score = b.score; a.idx = b.idx; a.cnt = b.cnt; a.sum = 0;
            for (int i = 0; i < (int)(blockDim.x >> 5); ++i) a.sum += s_sumw[i];
            partials[blockIdx.x] = a;
        }
        if (!mgpu_barrier<CLUSTER>(bar_counter, &s_epoch, nblocks, &st->abort_flag, &s_flag)) return;
        Best t{-1.0e308, 0x7fffffff, 0u};
        unsigned long long live_now = 0;
        for (unsigned int i = threadIdx.x; i < nblocks; i += blockDim.x) {
            Best o;
            o.score = __ldcg(&partials[i].score);
            o.idx = __ldcg(&partials[i].idx);
            o.cnt = __ldcg(&partials[i].cnt);
            live_now += __ldcg(&partials[i].sum);
            t = best_of(t, o);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) live_now += __shfl_xor_sync(0xffffffffu, live_now, o);
        __syncthreads();
        if (lane == 0) s_sumw[threadIdx.x >> 5] = live_now;
        b = block_best(t, s_red);
        live_now = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) live_now += s_sumw[i];
        if (p.S == 0 || b.score == 0.0) {                  // utmos/select.py:51-52
            stop = UTMOS_STOP_ZERO;
            break;
        }
        if (p.tail_budget && live_now <= p.tail_budget && b.cnt < p.tail_rows) {
            want_tail = 1;                                 // sparse enough: hand over to the replicated tail (mgpu.cu)
            break;
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            p.out_idx[step] = b.idx;
            p.out_new[step] = b.cnt;
            p.out_score[step] = b.score;
            if (p.dbg_time) p.out_time[step] = global_timer_ns();
            p.mask[b.idx] = 0;                             // utmos/select.py:100
        }
        step += 1;
        tot += b.cnt;
        if (tot >= m.global_V) {                           // utmos/select.py:110-112
            stop = UTMOS_STOP_ALL;
            break;
        }
        // ---- B: this rank's newly covered rows -> local delta.  A pick that covers very many rows (the first few:
        // common variants) is not subtracted bit by bit: the live mask is updated, then this rank's share of every
        // gain is recomputed by streaming its sample-major copy, and the delta is (new share - old share).
        const bool regain = p.cols && p.regain_rows && b.cnt >= p.regain_rows * (unsigned int)m.world;
        if (p.V > 0)
            for (long long c = warp0; c < nchunks; c += nwarps) cover_chunk(pd, b.idx, c, lane, !regain);
        if (!mgpu_barrier<CLUSTER>(bar_counter, &s_epoch, nblocks, &st->abort_flag, &s_flag)) return;
        if (regain) {
            const uint4 *lv = reinterpret_cast<const uint4 *>(p.live);
            const long long n4 = p.colPitchW / 4;
            // one warp per sample: 128-bit streaming loads of its column, shuffle reduction, no block-level sync
            for (long long s = warp0; s < p.S; s += nwarps) {
                const uint4 *col = reinterpret_cast<const uint4 *>(p.cols + s * p.colPitchW);
                unsigned int alive = 0;
                unsigned long long lo = 0, hi = 0;
                if (p.V > 0) {
                    for (long long i = lane; i < n4; i += 32) {
                        const uint4 c4 = ld_stream_u128(col + i);
                        const uint4 l4 = __ldcg(lv + i);
                        uint32_t w[4] = {c4.x & l4.x, c4.y & l4.y, c4.z & l4.z, c4.w & l4.w};
                        alive += __popc(w[0]) + __popc(w[1]) + __popc(w[2]) + __popc(w[3]);
                        if (p.af) {
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                uint32_t x = w[u];
                                while (x) {
                                    const long long r = (i * 4 + u) * 32 + (__ffs(x) - 1);
                                    x &= x - 1;
                                    lo += __ldg(p.q_lo + r);
                                    hi += __ldg(p.q_hi + r);
                                }
                            }
                        }
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    alive += __shfl_xor_sync(0xffffffffu, alive, o);
                    if (p.af) { lo += __shfl_xor_sync(0xffffffffu, lo, o); hi += __shfl_xor_sync(0xffffffffu, hi, o); }
                }
                if (lane == 0) {
                    m.delta_cnt[s] = alive - __ldcg(m.local_cnt + s);     // two's complement, <= 0
                    if (p.af) { m.delta_lo[s] = lo - __ldcg(m.local_lo + s); m.delta_hi[s] = hi - __ldcg(m.local_hi + s); }
                }
            }
            if (!mgpu_barrier<CLUSTER>(bar_counter, &s_epoch, nblocks, &st->abort_flag, &s_flag)) return;
        }
        // ---- C: push the delta into every peer's inbox (NVLink P2P stores), then publish the sequence number
        seq += 1;
        const size_t slot = ((size_t)(seq & 1) * m.world + m.rank) * (size_t)p.S;
        for (int q = 0; q < m.world; ++q) {
            if (q == m.rank) continue;
            unsigned int *dst = m.peer_inbox_cnt[q] + slot;
            for (long long i = gtid; i < p.S; i += nthreads) dst[i] = __ldcg(m.delta_cnt + i);
            if (p.af) {
                unsigned long long *dl = m.peer_inbox_lo[q] + slot, *dh = m.peer_inbox_hi[q] + slot;
                for (long long i = gtid; i < p.S; i += nthreads) { dl[i] = __ldcg(m.delta_lo + i); dh[i] = __ldcg(m.delta_hi + i); }
            }
        }
        __threadfence_system();
        if (!mgpu_barrier<CLUSTER>(bar_counter, &s_epoch, nblocks, &st->abort_flag, &s_flag)) return;
        if (blockIdx.x == 0 && threadIdx.x < m.world && (int)threadIdx.x != m.rank)
            st_release_sys_u64(m.peer_flags[threadIdx.x] + m.rank, seq);
        // ---- D: wait for every peer's delta of this step, apply the sum to the replica, clear the local delta
        if (threadIdx.x == 0) {
            int bad = 0;
            for (int q = 0; q < m.world && !bad; ++q) {
                if (q == m.rank) continue;
                long long spins = 0;
                while (ld_acquire_sys_u64(m.flags + q) < seq) {
                    if (++spins > (kSpinLimit >> 3)) { atomicExch(&st->abort_flag, 2u); bad = 1; break; }
                    if ((spins & 0xfff) == 0 && ld_acquire_u32(&st->abort_flag)) { bad = 1; break; }
                }
            }
            s_flag = bad;
        }
        __syncthreads();
        if (s_flag) return;
        const size_t base = (size_t)(seq & 1) * m.world * (size_t)p.S;
        for (long long i = gtid; i < p.S; i += nthreads) {
            unsigned int d = __ldcg(m.delta_cnt + i);
            unsigned long long dl = 0, dh = 0;
            if (p.af) { dl = __ldcg(m.delta_lo + i); dh = __ldcg(m.delta_hi + i); }
            for (int q = 0; q < m.world; ++q) {
                if (q == m.rank) continue;
                const size_t o = base + (size_t)q * p.S + i;
                d += __ldcv(m.inbox_cnt + o);
                if (p.af) { dl += __ldcv(m.inbox_lo + o); dh += __ldcv(m.inbox_hi + o); }
            }
            if (d) p.gain_cnt[i] += d;
            m.local_cnt[i] += __ldcg(m.delta_cnt + i);                   // this rank's share follows its own delta
            m.delta_cnt[i] = 0;
            if (p.af) {
                if (dl) p.gain_lo[i] += dl;
                if (dh) p.gain_hi[i] += dh;
                m.local_lo[i] += __ldcg(m.delta_lo + i);
                m.local_hi[i] += __ldcg(m.delta_hi + i);
                m.delta_lo[i] = 0;
                m.delta_hi[i] = 0;
            }
        }
        if (!mgpu_barrier<CLUSTER>(bar_counter, &s_epoch, nblocks, &st->abort_flag, &s_flag)) return;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->step = step;
        st->tot = tot;
        st->stop = stop;
        st->winner = -1;
        st->regain = 0;
        st->want_tail = want_tail;
        st->mgpu_seq = seq;
    }
}

__global__ void debug_scores_kernel(SelParams p, double *score_out)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= p.S) return;
    score_out[s] = p.af ? fixed_to_double(p.gain_lo[s], p.gain_hi[s], p.L, p.scale) : (double)p.gain_cnt[s];
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
int launch_transpose(cudaStream_t stream, const uint32_t *rows, long long V, int pitchW, int S, uint32_t *cols,
                     long long colPitchW, int *n_launch)
{
    if (V <= 0 || S <= 0) return UTMOS_OK;
    const int S32 = (S + 31) / 32 * 32;
    // colPitchW is a multiple of 8 words and covers ceil(V/32); one tile = 256 rows = 8 words
    const long long row_tiles = colPitchW / (kTRows / 32);
    const int col_tiles = (pitchW + kTWords - 1) / kTWords;
    if (row_tiles * col_tiles > 0x7fffffffll) { set_error("transpose: matrix too large"); return UTMOS_E_ARG; }
    static bool configured = false;
    static int flavour = 2;
    if (!configured) {
        UT_CUDA(cudaFuncSetAttribute(transpose_bits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTransposeSmem));
        UT_CUDA(cudaFuncSetAttribute(transpose_bits_reg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTransposeRegSmem));
        const char *env = getenv("UTMOS_B200_TRANSPOSE");          // A/B runs: 1 = shuffle flavour, default = register flavour
        if (env) flavour = atoi(env) == 1 ? 1 : 2;
        configured = true;
    }
    if (flavour == 2) {
        const int reg_col_tiles = (pitchW + kRWords - 1) / kRWords;
        if (row_tiles * reg_col_tiles > 0x7fffffffll) { set_error("transpose: matrix too large"); return UTMOS_E_ARG; }
        transpose_bits_reg_kernel<<<(unsigned)(row_tiles * reg_col_tiles), 256, kTransposeRegSmem, stream>>>(
            rows, V, pitchW, S32 / 32, cols, colPitchW, reg_col_tiles);
    } else {
        transpose_bits_kernel<<<(unsigned)(row_tiles * col_tiles), 256, kTransposeSmem, stream>>>(rows, V, pitchW, S32, cols,
                                                                                                 colPitchW, col_tiles);
    }
    *n_launch += 1;
    UT_CUDA(cudaGetLastError());
    return UTMOS_OK;
}

int launch_fixed_af(cudaStream_t stream, const double *af, long long V, int af_mode, int L, int scale,
                    unsigned long long *q_lo, unsigned long long *q_hi, SelState *st, int *n_launch)
{
    if (V <= 0) return UTMOS_OK;
    fixed_af_kernel<<<(unsigned)((V + 255) / 256), 256, 0, stream>>>(af, V, af_mode, L, scale, q_lo, q_hi, st);
    *n_launch += 1;
    UT_CUDA(cudaGetLastError());
    return UTMOS_OK;
}

int launch_live_init(cudaStream_t stream, uint32_t *live, long long colPitchW, long long V,
                     const unsigned long long *q_lo, const unsigned long long *q_hi, int af, int *n_launch)
{
    if (colPitchW <= 0) return UTMOS_OK;
    live_init_kernel<<<(unsigned)((colPitchW + 255) / 256), 256, 0, stream>>>(live, colPitchW, V, q_lo, q_hi, af);
    *n_launch += 1;
    UT_CUDA(cudaGetLastError());
    return UTMOS_OK;
}

// var_count / gain_cnt / gain_lo / gain_hi must be zeroed by the caller
int launch_gain_init(cudaStream_t stream, const SelParams &p, unsigned int *var_count, int *n_launch)
{
    if (p.V <= 0 || p.S <= 0) return UTMOS_OK;
    int what = 0;
    if (p.cols) {
        colpop_kernel<<<p.S, 256, 0, stream>>>(p.cols, p.live, p.colPitchW, p.S, var_count, p.gain_cnt);
        *n_launch += 1;
    } else {
        what |= 1 | 2;
    }
    if (p.af && p.cols) {
        const dim3 grid((unsigned)((p.colPitchW + kAfWords - 1) / kAfWords), (unsigned)((p.S + kAfSamples - 1) / kAfSamples));
        col_af_kernel<<<grid, 256, 0, stream>>>(p.cols, p.live, p.q_lo, p.q_hi, p.colPitchW, p.V, p.S, p.gain_lo, p.gain_hi);
        *n_launch += 1;
    } else if (p.af) {
        what |= 4;
    }
    if (!var_count) what &= ~1;                 // gains of a restored live mask only (utmos_select_import)
    if (what) {
        const long long warps_needed = p.V;
        long long blocks = (warps_needed + 7) / 8;
        if (blocks > 148 * 64) blocks = 148 * 64;
        row_gain_kernel<<<(unsigned)blocks, 256, 0, stream>>>(p, var_count, what);
        *n_launch += 1;
    }
    UT_CUDA(cudaGetLastError());
    return UTMOS_OK;
}

int launch_step_pair(cudaStream_t stream, const SelParams &p, int n_sms, int *n_launch)
{
    argmax_step_kernel<<<1, 1024, 0, stream>>>(p);
    const long long nchunks = (p.colPitchW + 31) / 32;
    long long blocks = (nchunks + 7) / 8;
    const long long cap = (long long)n_sms * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    cover_step_kernel<<<(unsigned)blocks, 256, 0, stream>>>(p);
    if (n_launch) *n_launch += 2;
    UT_CUDA(cudaGetLastError());
    return UTMOS_OK;
}

int persistent_grid(int device, int *grid_out, int *block_out)
{
    int n_sms = 0, per_sm = 0, coop = 0;
    UT_CUDA(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, device));
    UT_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
    if (!coop) { set_error("device does not support cooperative launch"); return UTMOS_E_NOGPU; }
    UT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, select_persistent_kernel, 1024, 0));
    if (per_sm < 1) { set_error("persistent kernel does not fit on an SM"); return UTMOS_E_CUDA; }
    *grid_out = n_sms;                // one CTA per SM
    *block_out = 1024;
    return UTMOS_OK;
}

int launch_persistent(cudaStream_t stream, const SelParams &p, int grid, int block, unsigned int *bar_counter,
                      ArgPartial *partials, int *n_launch)
{
    UT_CUDA(cudaMemsetAsync(bar_counter, 0, sizeof(unsigned int), stream));
    SelParams pp = p;
    void *args[] = {&pp, &bar_counter, &partials};
    UT_CUDA(cudaLaunchCooperativeKernel((void *)select_persistent_kernel, dim3(grid), dim3(block), args, 0, stream));
    *n_launch += 1;
    return UTMOS_OK;
}


// Cluster flavour: returns UTMOS_OK and *cluster_out = 0 when the state does not fit / clusters are unavailable.
static int cluster_layout(const SelParams &p, int CL, ClusterCfg *cfg, size_t *smem_bytes)
{
    cfg->gains_l2 = p.dsmem_gains ? 0 : 1;
    const int per = ((p.S + CL - 1) / CL + 31) / 32 * 32;
    const long long liveW = ((p.colPitchW + CL - 1) / CL + 31) / 32 * 32;
    if (liveW > (1 << 20)) return 0;
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = (off + bytes + 15) / 16 * 16; return (int)o; };
    take((size_t)per * 4);                                  // cnt at offset 0
    cfg->off_lo = take(p.af ? (size_t)per * 8 : 0);
    cfg->off_hi = take(p.af ? (size_t)per * 8 : 0);
    cfg->off_w = take(p.weights ? (size_t)per * 8 : 0);
    cfg->off_mask = take((size_t)per);
    cfg->off_live = take((size_t)liveW * 4);
    cfg->off_part = take(sizeof(ArgPartial) * 16);
    int G = 1;
    while (G < 32 && (p.pitchW / 4 + G - 1) / G > 8) G <<= 1;
    cfg->lanes_per_row = G;
    cfg->per = per;
    cfg->liveW = (int)liveW;
    *smem_bytes = off;
    return 1;
}

int cluster_plan(const SelParams &p, int *cluster_out)
{
    *cluster_out = 0;
    static int configured = 0;
    if (!configured) {
        UT_CUDA(cudaFuncSetAttribute(select_cluster_kernel<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        UT_CUDA(cudaFuncSetAttribute(select_cluster_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        configured = 1;
    }
    const int tries[2] = {16, 8};
    for (int t = 0; t < 2; ++t) {
        const int CL = tries[t];
        ClusterCfg cfg;
        size_t smem = 0;
        if (!cluster_layout(p, CL, &cfg, &smem)) continue;
        if (smem > 227 * 1024 - 1024) continue;
        if (cudaFuncSetAttribute(select_cluster_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
            cudaSuccess) { cudaGetLastError(); continue; }
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(CL);
        lc.blockDim = dim3(1024);
        lc.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        lc.attrs = attr;
        lc.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, select_cluster_kernel<false>, &lc) != cudaSuccess) { cudaGetLastError(); continue; }
        if (n >= 1) { *cluster_out = CL; return UTMOS_OK; }
    }
    return UTMOS_OK;
}

int launch_cluster(cudaStream_t stream, const SelParams &p, int CL, int *n_launch, uint32_t *newmask)
{
    ClusterCfg cfg;
    size_t smem = 0;
    if (!cluster_layout(p, CL, &cfg, &smem)) { set_error("cluster layout failed"); return UTMOS_E_ARG; }
    cfg.newmask = newmask;
    auto kernel = newmask ? select_cluster_kernel<true> : select_cluster_kernel<false>;
    UT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(CL);
    lc.blockDim = dim3(1024);
    lc.dynamicSmemBytes = smem;
    lc.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    lc.attrs = attr;
    lc.numAttrs = 1;
    UT_CUDA(cudaLaunchKernelEx(&lc, kernel, p, cfg));
    *n_launch += 1;
    return UTMOS_OK;
}

int launch_cover_decrement(cudaStream_t stream, const SelParams &p, const uint32_t *newmask, int *n_launch)
{
    if (p.S <= 0 || p.V <= 0 || !newmask || p.af) { set_error("cover_decrement: count mode with a newmask buffer only"); return UTMOS_E_ARG; }
    const int col_passes = (p.pitchW + 255) / 256;
    long long gx = std::max(1, 2 * 148 / col_passes);                       // about two CTAs per SM in all
    long long wpc = (p.colPitchW + gx - 1) / gx;
    wpc = std::min<long long>((wpc + kDecBatchWords - 1) / kDecBatchWords * kDecBatchWords, kDecMaxWords);
    gx = (p.colPitchW + wpc - 1) / wpc;
    if (gx > 0x7fffffffll) { set_error("cover_decrement: matrix too large"); return UTMOS_E_ARG; }
    const size_t smem = (size_t)std::min(256, p.pitchW) * 32 * sizeof(unsigned int);
    cover_decrement_kernel<<<dim3((unsigned)gx, (unsigned)col_passes), 256, smem, stream>>>(p, newmask, (int)wpc);
    *n_launch += 1;
    UT_CUDA(cudaGetLastError());
    return UTMOS_OK;
}

int launch_regain(cudaStream_t stream, const SelParams &p, int *n_launch)
{
    if (!p.cols || p.S <= 0) return UTMOS_OK;
    regain_kernel<<<p.S, 256, 0, stream>>>(p);
    *n_launch += 1;
    UT_CUDA(cudaGetLastError());
    return UTMOS_OK;
}



// grid_out < 0: |grid_out| CTAs as ONE thread-block cluster (hardware barrier); > 0: cooperative grid
int mgpu_grid(int device, int *grid_out, int *block_out)
{
    int n_sms = 0, per_sm = 0, coop = 0;
    UT_CUDA(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, device));
    UT_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
    *block_out = 1024;
    // UTMOS_B200_MGPU_GRID: n > 0 = cooperative grid of n CTAs, n < 0 = ONE cluster of |n| CTAs (hardware barrier).
    // Measured on 2 x B200 (1kGP shape): the 16-CTA cluster is slower than a 64-CTA grid -- the heavy steps need the
    // SMs (streaming recompute, retiring rows), the barrier latency is not what bounds them -- so the grid is the default.
    const char *env = getenv("UTMOS_B200_MGPU_GRID");
    const int want = env ? atoi(env) : 0;
    if (want < 0 && cudaFuncSetAttribute(select_mgpu_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
        const int tries[2] = {-want > 16 ? 16 : -want, 8};
        for (int t = 0; t < 2; ++t) {
            cudaLaunchConfig_t lc = {};
            lc.gridDim = dim3(tries[t]);
            lc.blockDim = dim3(1024);
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = tries[t]; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            lc.attrs = attr;
            lc.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, select_mgpu_kernel<true>, &lc) == cudaSuccess && n >= 1) {
                *grid_out = -tries[t];
                return UTMOS_OK;
            }
            cudaGetLastError();
        }
    }
    cudaGetLastError();
    if (!coop) { set_error("device does not support cooperative launch"); return UTMOS_E_NOGPU; }
    UT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, select_mgpu_kernel<false>, 1024, 0));
    if (per_sm < 1) { set_error("multi-GPU kernel does not fit on an SM"); return UTMOS_E_CUDA; }
    *grid_out = n_sms < 96 ? n_sms : 96;      // measured on 2 x B200: 32 CTAs 18.4 ms, 64: 16.7, 96: 15.5, 148: 15.8 per selection
    if (want > 0) *grid_out = want < n_sms * per_sm ? want : n_sms * per_sm;
    return UTMOS_OK;
}

int launch_mgpu(cudaStream_t stream, const SelParams &p, const MgpuParams &m, int grid, int block,
                unsigned int *bar_counter, ArgPartial *partials, int *n_launch)
{
    SelParams pp = p;
    MgpuParams mm = m;
    if (grid < 0) {
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(-grid);
        lc.blockDim = dim3(block);
        lc.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = -grid; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        lc.attrs = attr;
        lc.numAttrs = 1;
        UT_CUDA(cudaLaunchKernelEx(&lc, select_mgpu_kernel<true>, pp, mm, bar_counter, partials));
    } else {
        UT_CUDA(cudaMemsetAsync(bar_counter, 0, sizeof(unsigned int), stream));
        void *args[] = {&pp, &mm, &bar_counter, &partials};
        UT_CUDA(cudaLaunchCooperativeKernel((void *)select_mgpu_kernel<false>, dim3(grid), dim3(block), args, 0, stream));
    }
    *n_launch += 1;
    return UTMOS_OK;
}

int launch_debug_scores(cudaStream_t stream, const SelParams &p, double *score_out, int *n_launch)
{
    if (p.S <= 0) return UTMOS_OK;
    debug_scores_kernel<<<(p.S + 255) / 256, 256, 0, stream>>>(p, score_out);
    *n_launch += 1;
    UT_CUDA(cudaGetLastError());
    return UTMOS_OK;
}

}  // namespace utmos
