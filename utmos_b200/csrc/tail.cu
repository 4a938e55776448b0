// tail.cu -- the latency-bound tail of the greedy loop, run by ONE CTA out of shared memory.
//
// After the first picks have swallowed the common variants (head: cluster / grid-wide kernel + regain), the
// part of the matrix that still scores is very sparse: a few percent of the set bits, spread over rows that
// mostly have a handful of carriers.  A greedy step then needs (a) the argmax over S gains and (b) for every
// row the pick newly covers, the row's other carriers.  Both fit a single SM if the data is laid out for it:
//
//   edge lists   for every sample s: one 16-byte entry per live row r that carries s, holding r and up to
//                five OTHER carriers of r inline (uint16).  Rows with more than six carriers keep their
//                carrier list once in a side pool (uint16) and the entry points at it.  AF flavours
//                append the row's fixed-point AF (two uint64 limbs) -> 32-byte entries.
//   shared memory  gains (count + AF limbs), mask, weights, the list directory and the live bitmask.
//
// A step = argmax (REDUX-based, first-index tie-break) -> stream the winner's list (one dependent global
// read, contiguous) -> atomicAnd on the live bit decides "newly covered" -> shared-memory atomics on the
// gains of the inline carriers.  No inter-SM synchronisation, no second dependent global access.
// Every list is walked once (a sample is picked once); entries of rows covered meanwhile are skipped by the
// live bit.  When at most half of the listed entries are still live the kernel returns and the lists are
// re-compacted by a streaming filter pass, so dead entries never dominate.
//
// Reference semantics: utmos/select.py:24-53 (scores, mask, weights, argmax, zero stop) and :91-112.
#include "common.cuh"

namespace utmos {

namespace {

constexpr unsigned int kPooled = 0xffffu;      // entry.y low half: carriers live in the pool at [z, z + w)
constexpr int kInline = 5;

// ------------------------------------------------------------------------------------------------
// small kernels around the lists
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) sum_gains_kernel(SelParams p)
{
    __shared__ unsigned long long s_part[32];
    unsigned long long acc = 0;
    for (int s = threadIdx.x; s < p.S; s += blockDim.x) acc += p.gain_cnt[s];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int i = 0; i < 32; ++i) t += s_part[i];
        p.st->live_bits = t;
    }
}

// list_len[s] = gain_cnt[s]; list_off = exclusive scan (single CTA, running carry); cursor[s] = 0
__global__ void __launch_bounds__(1024) list_offsets_kernel(const unsigned int *gain_cnt, int S, unsigned int *list_off,
                                                            unsigned int *list_len, unsigned int *cursor)
{
    __shared__ unsigned int s_warp[32];
    __shared__ unsigned int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < S; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const unsigned int v = i < S ? gain_cnt[i] : 0u;
        unsigned int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const unsigned int w = s_warp[lane];
            unsigned int wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            s_warp[lane] = wi - w;
        }
        __syncthreads();
        const unsigned int carry = s_carry;
        if (i < S) {
            list_off[i] = carry + s_warp[warp] + incl - v;
            list_len[i] = v;
            if (cursor) cursor[i] = 0;
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = carry + s_warp[warp] + incl;
        __syncthreads();
    }
}

__device__ __forceinline__ uint4 pack_entry(unsigned int r, unsigned int n, const unsigned short *c)
{
    return make_uint4(r, n | ((unsigned int)c[0] << 16), (unsigned int)c[1] | ((unsigned int)c[2] << 16),
                      (unsigned int)c[3] | ((unsigned int)c[4] << 16));
}

// First compaction, from the bit matrix: one warp per live row.  A row with k carriers yields k entries, one
// in the list of each carrier.  ESTRIDE = 1 (count) or 2 (AF flavours: second uint4 = fixed-point AF limbs).
template <int ESTRIDE>
__global__ void __launch_bounds__(256) build_edges_kernel(SelParams p, uint4 *lists, const unsigned int *list_off,
                                                          unsigned int *cursor, unsigned short *pool,
                                                          unsigned int *pool_cursor)
{
    __shared__ unsigned short s_car[8][8];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = warp0; r < p.V; r += nwarps) {
        if (!((p.live[r >> 5] >> (r & 31)) & 1u)) continue;
        const uint32_t *row = p.rows + r * p.pitchW;
        int mine = 0;
        for (int k = lane; k < p.nW; k += 32) mine += __popc(__ldg(row + k));
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        uint4 tailq = make_uint4(0u, 0u, 0u, 0u);
        if (ESTRIDE == 2) {
            const unsigned long long ql = p.q_lo[r], qh = p.q_hi[r];
            tailq = make_uint4((unsigned int)ql, (unsigned int)(ql >> 32), (unsigned int)qh, (unsigned int)(qh >> 32));
        }
        if (total - 1 <= kInline) {
            int pos = incl - mine;
            for (int k = lane; k < p.nW; k += 32) {
                uint32_t x = __ldg(row + k);
                while (x) {
                    s_car[wib][pos++] = (unsigned short)((k << 5) + (__ffs(x) - 1));
                    x &= x - 1;
                }
            }
            __syncwarp();
            if (lane < total) {
                unsigned short others[kInline] = {0, 0, 0, 0, 0};
                int m = 0;
                for (int j = 0; j < total; ++j)
                    if (j != lane) others[m++] = s_car[wib][j];
                const unsigned int s = s_car[wib][lane];
                const unsigned int slot = atomicAdd(cursor + s, 1u);
                uint4 *dst = lists + ((size_t)list_off[s] + slot) * ESTRIDE;
                dst[0] = pack_entry((unsigned int)r, (unsigned int)(total - 1), others);
                if (ESTRIDE == 2) dst[1] = tailq;
            }
            __syncwarp();
        } else {
            // carriers of this row go to the pool once; every carrier's entry points at them
            unsigned int base = 0;
            if (lane == 0) base = atomicAdd(pool_cursor, (unsigned int)total);
            base = __shfl_sync(0xffffffffu, base, 0);
            int pos = incl - mine;
            for (int k = lane; k < p.nW; k += 32) {
                uint32_t x = __ldg(row + k);
                while (x) {
                    const unsigned int s = (unsigned int)((k << 5) + (__ffs(x) - 1));
                    x &= x - 1;
                    pool[base + pos++] = (unsigned short)s;
                    const unsigned int slot = atomicAdd(cursor + s, 1u);
                    uint4 *dst = lists + ((size_t)list_off[s] + slot) * ESTRIDE;
                    dst[0] = make_uint4((unsigned int)r, kPooled, base, (unsigned int)total);
                    if (ESTRIDE == 2) dst[1] = tailq;
                }
            }
        }
    }
}

// Re-compaction: one CTA per sample copies the entries whose row is still live (streaming filter).
template <int ESTRIDE>
__global__ void __launch_bounds__(256) filter_edges_kernel(const uint32_t *__restrict__ live,
                                                           const uint4 *__restrict__ old_lists,
                                                           const unsigned int *__restrict__ old_off,
                                                           const unsigned int *__restrict__ old_len,
                                                           uint4 *__restrict__ new_lists,
                                                           const unsigned int *__restrict__ new_off)
{
    __shared__ unsigned int s_cursor;
    const int s = blockIdx.x;
    if (threadIdx.x == 0) s_cursor = 0;
    __syncthreads();
    const uint4 *src = old_lists + (size_t)old_off[s] * ESTRIDE;
    uint4 *dst = new_lists + (size_t)new_off[s] * ESTRIDE;
    const int n = (int)old_len[s];
    const int lane = threadIdx.x & 31;
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        uint4 e = make_uint4(0u, 0u, 0u, 0u), e2 = e;
        bool keep = false;
        if (i < n) {
            e = __ldg(src + (size_t)i * ESTRIDE);
            keep = (__ldcg(live + (e.x >> 5)) >> (e.x & 31)) & 1u;
            if (keep && ESTRIDE == 2) e2 = __ldg(src + (size_t)i * ESTRIDE + 1);
        }
        const unsigned int m = __ballot_sync(0xffffffffu, keep);
        if (m) {
            const int leader = __ffs(m) - 1;
            unsigned int pos0 = 0;
            if (lane == leader) pos0 = atomicAdd(&s_cursor, (unsigned int)__popc(m));
            pos0 = __shfl_sync(0xffffffffu, pos0, leader);
            if (keep) {
                const unsigned int pos = pos0 + __popc(m & ((1u << lane) - 1u));
                dst[(size_t)pos * ESTRIDE] = e;
                if (ESTRIDE == 2) dst[(size_t)pos * ESTRIDE + 1] = e2;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// argmax with REDUX: scores become order-preserving 64-bit keys; max(hi) -> max(lo | hi==max) -> min(idx)
// ------------------------------------------------------------------------------------------------
struct Cand {
    unsigned int hi, lo;     // order-preserving key of the score
    int idx;
    unsigned int cnt;
};

__device__ __forceinline__ unsigned long long score_key(double s)
{
    s += 0.0;                                   // -0.0 -> +0.0 so that both zeros tie (np.argmax sees them equal)
    const unsigned long long b = (unsigned long long)__double_as_longlong(s);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__device__ __forceinline__ double key_score(unsigned int hi, unsigned int lo)
{
    const unsigned long long k = ((unsigned long long)hi << 32) | lo;
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

__device__ __forceinline__ bool cand_better(const Cand &a, const Cand &b)
{
    return a.hi > b.hi || (a.hi == b.hi && (a.lo > b.lo || (a.lo == b.lo && a.idx < b.idx)));
}

// all lanes return the warp's best candidate
__device__ __forceinline__ Cand warp_argmax(Cand c)
{
    const unsigned int mh = __reduce_max_sync(0xffffffffu, c.hi);
    const unsigned int ml = __reduce_max_sync(0xffffffffu, c.hi == mh ? c.lo : 0u);
    const bool top = c.hi == mh && c.lo == ml;
    const int mi = (int)__reduce_min_sync(0xffffffffu, top ? (unsigned int)c.idx : 0x7fffffffu);
    const unsigned int who = __ballot_sync(0xffffffffu, top && c.idx == mi);
    Cand out;
    out.hi = mh;
    out.lo = ml;
    out.idx = mi;
    out.cnt = __shfl_sync(0xffffffffu, c.cnt, __ffs(who) - 1);
    return out;
}

struct TailCfg {
    int off_lo, off_hi, off_w, off_mask, off_loff, off_llen, off_live, off_queue;   // byte offsets, counts at 0
    int qcap;            // pooled-row queue entries (3 words each: pool base, carriers, row)
    int live_words;      // > 0: live mask held in shared memory
    int lanes_per_row;   // power of two: lanes sharing one overflow row
    unsigned int min_recompact;   // do not bother re-compacting below this many live entries
};

template <int ESTRIDE>
__global__ void __launch_bounds__(1024, 1) select_tail_kernel(SelParams p, TailCfg cfg, unsigned long long lists_total)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ Cand s_red[32];
    __shared__ unsigned long long s_sum[32];
    __shared__ unsigned int s_qn;
    constexpr bool AF = ESTRIDE == 2;
    unsigned int *s_cnt = reinterpret_cast<unsigned int *>(smem);
    unsigned long long *s_lo = reinterpret_cast<unsigned long long *>(smem + cfg.off_lo);
    unsigned long long *s_hi = reinterpret_cast<unsigned long long *>(smem + cfg.off_hi);
    double *s_w = reinterpret_cast<double *>(smem + cfg.off_w);
    uint8_t *s_mask = smem + cfg.off_mask;
    unsigned int *s_loff = reinterpret_cast<unsigned int *>(smem + cfg.off_loff);
    unsigned int *s_llen = reinterpret_cast<unsigned int *>(smem + cfg.off_llen);
    uint32_t *s_live = reinterpret_cast<uint32_t *>(smem + cfg.off_live);
    unsigned int *s_queue = reinterpret_cast<unsigned int *>(smem + cfg.off_queue);
    const unsigned short *pool = p.pool;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool has_w = p.weights != nullptr;
    const bool live_smem = cfg.live_words > 0;
    SelState *st = p.st;

    for (int i = tid; i < p.S; i += blockDim.x) {
        s_cnt[i] = p.gain_cnt[i];
        s_mask[i] = p.mask[i];
        s_loff[i] = p.list_off[i];
        s_llen[i] = p.list_len[i];
        if (AF) { s_lo[i] = p.gain_lo[i]; s_hi[i] = p.gain_hi[i]; }
        if (has_w) s_w[i] = p.weights[i];
    }
    for (int i = tid; i < cfg.live_words; i += blockDim.x) s_live[i] = p.live[i];
    long long step = st->step, tot = st->tot;
    const long long limit = st->limit;
    int stop = st->stop;
    int recompact = 0;
    int since_check = 0;
    if (tid == 0) s_qn = 0;
    __syncthreads();
    long long t_arg = 0, t_walk = 0, t_ret = 0, t_mark = clock64();
#define UT_TICK(acc) do { const long long now__ = clock64(); acc += now__ - t_mark; t_mark = now__; } while (0)

    while (stop == 0 && step < limit) {
        // ---- every 64 steps: how many list entries are still live?  (sum of the gains)
        if (++since_check >= 64) {
            since_check = 0;
            unsigned long long acc = 0;
            for (int i = tid; i < p.S; i += blockDim.x) acc += s_cnt[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) s_sum[warp] = acc;
            __syncthreads();
            unsigned long long live_now = s_sum[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) live_now += __shfl_xor_sync(0xffffffffu, live_now, o);
            __syncthreads();
            if (live_now >= cfg.min_recompact && live_now * 2 <= lists_total && limit - step > 128) {
                recompact = 1;
                break;
            }
        }
        // ---- argmax over all samples (shared memory, np.argmax order)
        Cand b{0u, 0u, 0x7fffffff, 0u};
        for (int i = tid; i < p.S; i += blockDim.x) {
            const unsigned int c = s_cnt[i];
            double g = 0.0;
            if (s_mask[i] == 1) {
                g = AF ? fixed_to_double(s_lo[i], s_hi[i], p.L, p.scale) : (double)c;
                if (has_w) g *= s_w[i];
            }
            const unsigned long long k = score_key(g);
            Cand c2{(unsigned int)(k >> 32), (unsigned int)k, i, c};
            if (cand_better(c2, b)) b = c2;
        }
        b = warp_argmax(b);
        if (lane == 0) s_red[warp] = b;
        __syncthreads();
        const Cand mine = s_red[lane];
        b = warp_argmax(mine);
        const double best_score = key_score(b.hi, b.lo);
        if (p.S == 0 || best_score == 0.0) {              // utmos/select.py:51-52
            stop = UTMOS_STOP_ZERO;
            break;
        }
        Cand t2 = mine;
        if (mine.idx == b.idx) { t2.hi = 0u; t2.lo = 0u; t2.idx = 0x7fffffff; }
        t2 = warp_argmax(t2);                             // best of the other warps' winners: a likely next pick
        if (tid == 0) {
            p.out_idx[step] = b.idx;
            p.out_new[step] = b.cnt;
            p.out_score[step] = best_score;
            if (p.dbg_time) p.out_time[step] = global_timer_ns();
            s_mask[b.idx] = 0;                            // utmos/select.py:100
        }
        step += 1;
        tot += b.cnt;
        if (tot >= p.V) {                                 // utmos/select.py:110-112
            stop = UTMOS_STOP_ALL;
            break;
        }
        UT_TICK(t_arg);
        // ---- stream the winner's list; a live bit that we clear marks a newly covered row
        const uint4 *lst = p.lists + (size_t)s_loff[b.idx] * ESTRIDE;
        const int len = (int)s_llen[b.idx];
        if (t2.idx != 0x7fffffff && key_score(t2.hi, t2.lo) > 0.0) {   // warm L2 with the runner-up's list
            const uint4 *l2 = p.lists + (size_t)s_loff[t2.idx] * ESTRIDE;
            const int lines = ((int)s_llen[t2.idx] * ESTRIDE + 7) >> 3;    // 128-byte lines
            for (int i = tid; i < lines; i += blockDim.x) prefetch_l2(l2 + (size_t)i * 8);
        }
        for (int base = 0; base < len; base += 4 * (int)blockDim.x) {
            uint4 e[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = base + u * (int)blockDim.x + tid;
                e[u] = i < len ? __ldg(lst + (size_t)i * ESTRIDE) : make_uint4(0xffffffffu, 0u, 0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned int r = e[u].x;
                bool fresh = false;
                if (r != 0xffffffffu) {
                    const uint32_t bit = 1u << (r & 31);
                    const uint32_t old = live_smem ? atomicAnd(s_live + (r >> 5), ~bit) : atomicAnd(p.live + (r >> 5), ~bit);
                    fresh = (old & bit) != 0;
                }
                const unsigned int n = e[u].y & 0xffffu;
                if (fresh && n != kPooled) {
                    unsigned long long nl = 0, nh = 0;
                    if (AF) {
                        const int i = base + u * (int)blockDim.x + tid;
                        const uint4 qv = __ldg(lst + (size_t)i * ESTRIDE + 1);
                        nl = 0ull - (((unsigned long long)qv.y << 32) | qv.x);
                        nh = 0ull - (((unsigned long long)qv.w << 32) | qv.z);
                        atomicAdd(s_lo + b.idx, nl);     // the pick's own gain (keeps sum(gains) == live entries)
                        atomicAdd(s_hi + b.idx, nh);
                    }
                    atomicAdd(s_cnt + b.idx, 0xffffffffu);
                    const unsigned int c[kInline] = {e[u].y >> 16, e[u].z & 0xffffu, e[u].z >> 16, e[u].w & 0xffffu,
                                                     e[u].w >> 16};
#pragma unroll
                    for (int j = 0; j < kInline; ++j) {
                        if (j < (int)n) {
                            atomicAdd(s_cnt + c[j], 0xffffffffu);
                            if (AF) { atomicAdd(s_lo + c[j], nl); atomicAdd(s_hi + c[j], nh); }
                        }
                    }
                }
                // rows with many carriers: their carrier list is in the pool -> warp-aggregated queue append
                const bool big = fresh && n == kPooled;
                const unsigned int m = __ballot_sync(0xffffffffu, big);
                if (m) {
                    const int leader = __ffs(m) - 1;
                    unsigned int pos0 = 0;
                    if (lane == leader) pos0 = atomicAdd(&s_qn, (unsigned int)__popc(m));
                    pos0 = __shfl_sync(0xffffffffu, pos0, leader);
                    if (big) {
                        const unsigned int pos = pos0 + __popc(m & ((1u << lane) - 1u));
                        if ((int)pos < cfg.qcap) {
                            s_queue[3 * pos] = e[u].z;
                            s_queue[3 * pos + 1] = e[u].w;
                            s_queue[3 * pos + 2] = r;
                        } else {
                            // queue full (exotic inputs only): this thread retires the row alone
                            unsigned long long nl = 0, nh = 0;
                            if (AF) { nl = 0ull - p.q_lo[r]; nh = 0ull - p.q_hi[r]; }
                            for (unsigned int k = 0; k < e[u].w; ++k) {
                                const unsigned int s = pool[e[u].z + k];
                                atomicAdd(s_cnt + s, 0xffffffffu);
                                if (AF) { atomicAdd(s_lo + s, nl); atomicAdd(s_hi + s, nh); }
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
        UT_TICK(t_walk);
        // ---- pooled rows: 8 lanes per row stream its carrier list (uint16) and decrement every carrier
        const int qn = min((int)s_qn, cfg.qcap);
        if (qn > 0) {
            const int sub = lane & 7, slot = lane >> 3;
            for (int q0 = warp * 4; q0 < qn; q0 += 32 * 4) {
                const int q = q0 + slot;
                if (q < qn) {
                    const unsigned int pbase = s_queue[3 * q], cnt = s_queue[3 * q + 1];
                    unsigned long long nl = 0, nh = 0;
                    if (AF) { const unsigned int r = s_queue[3 * q + 2]; nl = 0ull - p.q_lo[r]; nh = 0ull - p.q_hi[r]; }
                    for (unsigned int k0 = 0; k0 < cnt; k0 += 32) {
                        unsigned int c4[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const unsigned int k = k0 + u * 8 + sub;
                            c4[u] = k < cnt ? (unsigned int)__ldg(pool + pbase + k) : 0xffffffffu;
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (c4[u] != 0xffffffffu) {
                                atomicAdd(s_cnt + c4[u], 0xffffffffu);
                                if (AF) { atomicAdd(s_lo + c4[u], nl); atomicAdd(s_hi + c4[u], nh); }
                            }
                        }
                    }
                }
            }
            __syncthreads();
            if (tid == 0) s_qn = 0;                       // ordered before the next walk by the argmax barrier
        }
        UT_TICK(t_ret);
    }
#undef UT_TICK
    if (tid == 0 && p.dbg) { p.dbg[8] += t_arg; p.dbg[9] += t_walk; p.dbg[10] += t_ret; p.dbg[11] += 1; }

    __syncthreads();
    for (int i = tid; i < p.S; i += blockDim.x) {
        p.gain_cnt[i] = s_cnt[i];
        p.mask[i] = s_mask[i];
        if (AF) { p.gain_lo[i] = s_lo[i]; p.gain_hi[i] = s_hi[i]; }
    }
    for (int i = tid; i < cfg.live_words; i += blockDim.x) p.live[i] = s_live[i];
    if (tid == 0) {
        st->step = step;
        st->tot = tot;
        st->stop = stop;
        st->winner = -1;
        st->regain = 0;
        st->recompact = recompact;
    }
}

int tail_layout(const SelParams &p, TailCfg *cfg, size_t *smem_bytes)
{
    if (p.S > 65535 || p.V >= 0xffffffffll) return 0;       // carriers are uint16, rows uint32 in the edge lists
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = (off + bytes + 15) / 16 * 16; return (int)o; };
    const size_t S = (size_t)p.S;
    take(S * 4);
    cfg->off_lo = take(p.af ? S * 8 : 0);
    cfg->off_hi = take(p.af ? S * 8 : 0);
    cfg->off_w = take(p.weights ? S * 8 : 0);
    cfg->off_mask = take(S);
    cfg->off_loff = take(S * 4);
    cfg->off_llen = take(S * 4);
    const size_t budget = 225 * 1024;
    if (off + 4096 > budget) return 0;
    const size_t live_bytes = (size_t)p.colPitchW * 4;
    cfg->live_words = 0;
    cfg->off_live = (int)off;
    if (off + live_bytes + 4096 <= budget) { cfg->off_live = take(live_bytes); cfg->live_words = (int)p.colPitchW; }
    size_t q = (budget - off) / 12;
    if (q > 4096) q = 4096;
    cfg->qcap = (int)q;
    cfg->off_queue = take(q * 12);
    int G = 1;
    while (G < 32 && (p.pitchW / 4 + G - 1) / G > 8) G <<= 1;
    cfg->lanes_per_row = G;
    cfg->min_recompact = 1u << 16;
    *smem_bytes = off;
    return 1;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
int launch_sum_gains(cudaStream_t stream, const SelParams &p, int *n_launch)
{
    sum_gains_kernel<<<1, 1024, 0, stream>>>(p);
    *n_launch += 1;
    UT_CUDA(cudaGetLastError());
    return UTMOS_OK;
}

int launch_build_lists(cudaStream_t stream, const SelParams &p, uint4 *lists, unsigned int *list_off,
                       unsigned int *list_len, unsigned int *cursor, unsigned short *pool, unsigned int *pool_cursor,
                       int *n_launch)
{
    UT_CUDA(cudaMemsetAsync(pool_cursor, 0, 4, stream));
    list_offsets_kernel<<<1, 1024, 0, stream>>>(p.gain_cnt, p.S, list_off, list_len, cursor);
    long long blocks = (p.V + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    if (p.af) build_edges_kernel<2><<<(unsigned)blocks, 256, 0, stream>>>(p, lists, list_off, cursor, pool, pool_cursor);
    else build_edges_kernel<1><<<(unsigned)blocks, 256, 0, stream>>>(p, lists, list_off, cursor, pool, pool_cursor);
    *n_launch += 2;
    UT_CUDA(cudaGetLastError());
    return UTMOS_OK;
}

int launch_filter_lists(cudaStream_t stream, const SelParams &p, const uint4 *old_lists, const unsigned int *old_off,
                        const unsigned int *old_len, uint4 *new_lists, unsigned int *new_off, unsigned int *new_len,
                        int *n_launch)
{
    list_offsets_kernel<<<1, 1024, 0, stream>>>(p.gain_cnt, p.S, new_off, new_len, nullptr);
    if (p.af) filter_edges_kernel<2><<<p.S, 256, 0, stream>>>(p.live, old_lists, old_off, old_len, new_lists, new_off);
    else filter_edges_kernel<1><<<p.S, 256, 0, stream>>>(p.live, old_lists, old_off, old_len, new_lists, new_off);
    *n_launch += 2;
    UT_CUDA(cudaGetLastError());
    return UTMOS_OK;
}

int tail_plan(const SelParams &p, int *ok_out)
{
    TailCfg cfg;
    size_t smem = 0;
    *ok_out = p.cols != nullptr && p.S > 0 && tail_layout(p, &cfg, &smem);
    return UTMOS_OK;
}

int launch_tail(cudaStream_t stream, const SelParams &p, unsigned long long lists_total, int *n_launch)
{
    TailCfg cfg;
    size_t smem = 0;
    if (!tail_layout(p, &cfg, &smem)) { set_error("tail kernel: state does not fit in shared memory"); return UTMOS_E_ARG; }
    if (p.af) {
        UT_CUDA(cudaFuncSetAttribute(select_tail_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        select_tail_kernel<2><<<1, 1024, smem, stream>>>(p, cfg, lists_total);
    } else {
        UT_CUDA(cudaFuncSetAttribute(select_tail_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        select_tail_kernel<1><<<1, 1024, smem, stream>>>(p, cfg, lists_total);
    }
    *n_launch += 1;
    UT_CUDA(cudaGetLastError());
    return UTMOS_OK;
}

}  // namespace utmos
