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

#include "common.cuh"

namespace utmos {

namespace {

// ------------------------------------------------------------------------------------------------
// K2b: bit-matrix transpose.  CTA tile = 128 rows x 32 words (1024 samples) staged through shared
// memory; each warp transposes 32x32 bit blocks with ballots; output is written as 16-byte pieces
// (128 row-bits) per sample.
// ------------------------------------------------------------------------------------------------
constexpr int kTRows = 128;

__global__ void __launch_bounds__(256) transpose_bits_kernel(const uint32_t *__restrict__ rows, long long V,
                                                             int pitchW, int S32, uint32_t *__restrict__ cols,
                                                             long long colPitchW)
{
    __shared__ uint32_t s_in[kTRows][33];
    __shared__ uint32_t s_out[1024][5];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long r0 = (long long)blockIdx.x * kTRows;
    const int w0 = blockIdx.y * 32;

    for (int i = warp; i < kTRows; i += 8) {
        const long long r = r0 + i;
        const int w = w0 + lane;
        s_in[i][lane] = (r < V && w < pitchW) ? rows[r * pitchW + w] : 0u;
    }
    __syncthreads();
    // warp `warp` -> row group g = warp & 3 (32 rows), word columns c = (warp >> 2) * 16 .. +16
    const int g = warp & 3;
    const int c_begin = (warp >> 2) * 16;
    for (int c = c_begin; c < c_begin + 16; ++c) {
        const uint32_t x = s_in[g * 32 + lane][c];
        uint32_t mine = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const uint32_t b = __ballot_sync(0xffffffffu, (x >> j) & 1u);
            if (lane == j) mine = b;
        }
        s_out[c * 32 + lane][g] = mine;
    }
    __syncthreads();
    // 1024 samples x 4 words: thread t writes samples t, t+256, ... as one 16-byte store each
    for (int sl = threadIdx.x; sl < 1024; sl += 256) {
        const int s = w0 * 32 + sl;
        if (s >= S32) continue;
        const uint4 v = make_uint4(s_out[sl][0], s_out[sl][1], s_out[sl][2], s_out[sl][3]);
        *reinterpret_cast<uint4 *>(cols + (long long)s * colPitchW + (r0 >> 5)) = v;
    }
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
        var_count[s] = a;
        gain_cnt[s] = b;
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
__device__ __forceinline__ void cover_chunk(const SelParams &p, int b, long long chunk, int lane)
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
__global__ void __launch_bounds__(1024) argmax_step_kernel(SelParams p)
{
    __shared__ Best s_red[32];
    SelState *st = p.st;
    const bool idle = st->stop != 0 || st->step >= st->limit;
    __syncthreads();
    if (idle) {
        if (threadIdx.x == 0) st->winner = -1;
        return;
    }
    Best b = block_best(scan_best(p, 0, p.S), s_red);
    if (threadIdx.x == 0) {
        if (p.S == 0 || b.score == 0.0) {                 // utmos/select.py:51-52
            st->stop = UTMOS_STOP_ZERO;
            st->winner = -1;
        } else {
            const long long i = st->step;
            p.out_idx[i] = b.idx;
            p.out_new[i] = b.cnt;
            p.out_score[i] = b.score;
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
    if (b < 0) return;
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long nchunks = (p.colPitchW + 31) / 32;
    for (long long c = warp0; c < nchunks; c += nwarps) cover_chunk(p, b, c, lane);
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
    const int per = (p.S + (int)nblocks - 1) / (int)nblocks;
    const int my_begin = min(p.S, (int)blockIdx.x * per), my_end = min(p.S, my_begin + per);
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)nblocks * blockDim.x) >> 5;
    const long long nchunks = (p.colPitchW + 31) / 32;
    __syncthreads();

    while (stop == 0 && step < limit) {
        // ---- phase A: partial argmax over this CTA's slice of the samples
        Best b = block_best(scan_best(p, my_begin, my_end), s_red);
        if (threadIdx.x == 0) {
            ArgPartial a;
            a.score = b.score; a.idx = b.idx; a.cnt = b.cnt;
            partials[blockIdx.x] = a;
        }
        if (!grid_barrier(bar_counter, &s_epoch, nblocks, &st->abort_flag, &s_flag)) return;
        Best t{-1.0e308, 0x7fffffff, 0u};
        for (unsigned int i = threadIdx.x; i < nblocks; i += blockDim.x) {
            Best o;
            o.score = __ldcg(&partials[i].score);
            o.idx = __ldcg(&partials[i].idx);
            o.cnt = __ldcg(&partials[i].cnt);
            t = best_of(t, o);
        }
        b = block_best(t, s_red);
        if (p.S == 0 || b.score == 0.0) {                  // utmos/select.py:51-52
            stop = UTMOS_STOP_ZERO;
            break;
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            p.out_idx[step] = b.idx;
            p.out_new[step] = b.cnt;
            p.out_score[step] = b.score;
            p.mask[b.idx] = 0;                             // utmos/select.py:100
        }
        step += 1;
        tot += b.cnt;
        if (tot >= p.V) {                                  // utmos/select.py:110-112
            stop = UTMOS_STOP_ALL;
            break;
        }
        // ---- phase B: clear the newly covered rows, subtract them from every gain
        for (long long c = warp0; c < nchunks; c += nwarps) cover_chunk(p, b.idx, c, lane);
        if (!grid_barrier(bar_counter, &s_epoch, nblocks, &st->abort_flag, &s_flag)) return;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->step = step;
        st->tot = tot;
        st->stop = stop;
        st->winner = -1;
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
    // colPitchW is a multiple of 8 words and covers ceil(V/32); tiles of 128 rows = 4 words
    const long long row_tiles = colPitchW / 4;
    const int col_tiles = (pitchW + 31) / 32;
    if (row_tiles > 0x7fffffffll) { set_error("transpose: too many rows"); return UTMOS_E_ARG; }
    dim3 grid((unsigned)row_tiles, (unsigned)col_tiles);
    transpose_bits_kernel<<<grid, 256, 0, stream>>>(rows, V, pitchW, S32, cols, colPitchW);
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
    if (p.af) what |= 4;
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

int launch_debug_scores(cudaStream_t stream, const SelParams &p, double *score_out, int *n_launch)
{
    if (p.S <= 0) return UTMOS_OK;
    debug_scores_kernel<<<(p.S + 255) / 256, 256, 0, stream>>>(p, score_out);
    *n_launch += 1;
    UT_CUDA(cudaGetLastError());
    return UTMOS_OK;
}

}  // namespace utmos
