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
#include <cooperative_groups.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace utmos {

namespace cg = cooperative_groups;

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
    // the gains of the SELECTABLE samples: they are what the edge lists hold (carriers that can never be picked get no
    // entries and, once the lists exist, no decrements either -- their gains go stale, nobody reads them)
    for (int s = threadIdx.x; s < p.S; s += blockDim.x) acc += p.mask[s] == 1 ? p.gain_cnt[s] : 0u;
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

// list_len[s] = gain_cnt[s] for selectable samples (0 for the others: they are never walked); list_off = exclusive
// scan (single CTA, running carry); cursor[s] = 0
__global__ void __launch_bounds__(1024) list_offsets_kernel(const unsigned int *gain_cnt, const uint8_t *mask, int S,
                                                            unsigned int *list_off, unsigned int *list_len, unsigned int *cursor)
{
    __shared__ unsigned int s_warp[32];
    __shared__ unsigned int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < S; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const unsigned int v = (i < S && mask[i] == 1) ? gain_cnt[i] : 0u;
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

// First compaction, from the bit matrix: one warp per live row.  A row with k carriers yields k entries, one
// in the list of each carrier.  ESTRIDE = 1 (count) or 2 (AF flavours: second uint4 = fixed-point AF limbs).
// EdgeDst says where the entries go: single GPU = this context's buffers; multi-GPU = the SAME slots of every
// rank's merged buffers (peer pointers, NVLink stores), so all ranks end up with byte-identical lists.
template <int ESTRIDE, bool WIDE>
__global__ void __launch_bounds__(256) build_edges_kernel(SelParams p, EdgeDst d, unsigned int *cursor,
                                                          unsigned int *pool_cursor)
{
    constexpr int kInl = WIDE ? 2 : kInline;           // other carriers stored inline in an entry
    constexpr int kPerChunk = WIDE ? 4 : 8;            // pool elements per 16-byte chunk
    __shared__ unsigned int s_car[8][8];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const unsigned int pool_base = d.pool_base ? *d.pool_base : 0u;
    for (long long r = warp0; r < p.V; r += nwarps) {
        if (!((p.live[r >> 5] >> (r & 31)) & 1u)) continue;
        const uint32_t *row = p.rows + r * p.pitchW;
        const unsigned int rg = (unsigned int)(r + d.row_base);          // row id in the (merged) live mask
        // only carriers that can still be picked get entries (and are listed as other carriers): p.selw = bitmask of the
        // samples with mask == 1 at select_begin.  Excluded samples (utmos/select.py:168-175) never score, so nobody needs
        // their gains; on a run with --subset this halves the lists and the decrements of the tail.
        int mine = 0;
        for (int k = lane; k < p.nW; k += 32) mine += __popc(__ldg(row + k) & p.selw[k]);
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) continue;                                        // carried by excluded samples only
        uint4 tailq = make_uint4(0u, 0u, 0u, 0u);
        if (ESTRIDE == 2) {
            const unsigned long long ql = p.q_lo[r], qh = p.q_hi[r];
            tailq = make_uint4((unsigned int)ql, (unsigned int)(ql >> 32), (unsigned int)qh, (unsigned int)(qh >> 32));
        }
        if (total - 1 <= kInl) {
            int pos = incl - mine;
            for (int k = lane; k < p.nW; k += 32) {
                uint32_t x = __ldg(row + k) & p.selw[k];
                while (x) {
                    s_car[wib][pos++] = (unsigned int)((k << 5) + (__ffs(x) - 1));
                    x &= x - 1;
                }
            }
            __syncwarp();
            if (lane < total) {
                unsigned int others[kInline] = {0, 0, 0, 0, 0};
                int m = 0;
                for (int j = 0; j < total; ++j)
                    if (j != lane) others[m++] = s_car[wib][j];
                const unsigned int s = s_car[wib][lane];
                const size_t slot = ((size_t)d.slot_base[s] + atomicAdd(cursor + s, 1u)) * ESTRIDE;
                uint4 e0;
                if (WIDE) e0 = make_uint4(rg, (unsigned int)(total - 1), others[0], others[1]);
                else e0 = make_uint4(rg, (unsigned int)(total - 1) | (others[0] << 16), others[1] | (others[2] << 16),
                                     others[3] | (others[4] << 16));
                for (int q = 0; q < d.world; ++q) {
                    d.lists[q][slot] = e0;
                    if (ESTRIDE == 2) d.lists[q][slot + 1] = tailq;
                }
            }
            __syncwarp();
        } else {
            // carriers of this row go to the pool once; every carrier's entry points at them
            // (padded to whole 16-byte chunks with all-ones so the tail kernel can use aligned 128-bit copies)
            const int padded = (total + kPerChunk - 1) / kPerChunk * kPerChunk;
            unsigned int base = 0;
            if (lane == 0) base = pool_base + atomicAdd(pool_cursor, (unsigned int)padded);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (lane < padded - total) {
                for (int q = 0; q < d.world; ++q) {
                    if (WIDE) reinterpret_cast<unsigned int *>(d.pool[q])[base + total + lane] = 0xffffffffu;
                    else d.pool[q][base + total + lane] = (unsigned short)0xffffu;
                }
            }
            int pos = incl - mine;
            for (int k = lane; k < p.nW; k += 32) {
                uint32_t x = __ldg(row + k) & p.selw[k];
                while (x) {
                    const unsigned int s = (unsigned int)((k << 5) + (__ffs(x) - 1));
                    x &= x - 1;
                    const size_t slot = ((size_t)d.slot_base[s] + atomicAdd(cursor + s, 1u)) * ESTRIDE;
                    const uint4 e0 = make_uint4(rg, WIDE ? 0xffffffffu : kPooled, base, (unsigned int)total);
                    for (int q = 0; q < d.world; ++q) {
                        if (WIDE) reinterpret_cast<unsigned int *>(d.pool[q])[base + pos] = s;
                        else d.pool[q][base + pos] = (unsigned short)s;
                        d.lists[q][slot] = e0;
                        if (ESTRIDE == 2) d.lists[q][slot + 1] = tailq;
                    }
                    pos++;
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

// 64-bit add on shared memory from two native 32-bit atomics.  atomicAdd(unsigned long long *) on shared memory is a
// compare-and-swap loop on sm_100a (SASS: ATOMS.CAST.SPIN.64); the AF flavours retire two 64-bit limbs per decrement,
// so that loop bounded their tail.  The low word's atomic returns the old value, which tells exactly whether THIS add
// wrapped; the wrap is folded into the high word's add.  Sums are exact modulo 2^64 whatever the interleaving; readers
// only look after a barrier.
__device__ __forceinline__ void smem_add64(unsigned long long *p, unsigned long long v)
{
    unsigned int *w = reinterpret_cast<unsigned int *>(p);
    const unsigned int lo = (unsigned int)v, hi = (unsigned int)(v >> 32);
    const unsigned int old = atomicAdd(w, lo);
    const unsigned int carry = (old + lo) < old ? 1u : 0u;
    if (hi + carry) atomicAdd(w + 1, hi + carry);
}

struct TailCfg {
    int off_lo, off_hi, off_w, off_mask, off_loff, off_llen, off_live;   // byte offsets, counts at 0
    int live_words;      // > 0: live mask held in shared memory
    unsigned int min_recompact;   // do not bother re-compacting below this many live entries
    unsigned int single_rows;     // cluster flavour: hand over to the single-CTA flavour once a pick covers fewer rows
    uint32_t *live_priv;          // cluster flavour with the live mask in global memory: [CL][colPitchW] private copies
    int off_stage;                // staging area for the carrier lists of the rows a pick newly covers
    unsigned int stage_cap;       // ... capacity in chunks (16 bytes of carriers; AF flavours: + 16 bytes of limbs)
};

// integer argmax for count mode without weights: key = gain count of a selectable sample (0 otherwise);
// np.argmax order = larger key, then lower index.  All lanes return the warp's winner.
__device__ __forceinline__ uint2 warp_argmax_u32(unsigned int key, unsigned int idx)
{
    const unsigned int mk = __reduce_max_sync(0xffffffffu, key);
    const unsigned int mi = __reduce_min_sync(0xffffffffu, key == mk ? idx : 0x7fffffffu);
    return make_uint2(mk, mi);
}

// ESTRIDE 1: count entries; 2: AF flavours (second uint4 = fixed-point AF limbs of the row).
// FAST: count mode without weights -> integer keys (two REDUX per reduction level instead of the float64 path).
// WIDE: more than 65,535 samples -> 32-bit carriers (two inline per entry, four per pooled 16-byte chunk).
// CL: 1 = one CTA does everything.  CL > 1 = a thread-block cluster of CL CTAs, "owner computes": every CTA of the
//     cluster walks the same list and keeps its own copy of the live mask (so all CTAs see the same newly covered
//     rows without talking), but keeps only the state of the samples it owns (s % CL == rank, local index s / CL),
//     applies only their decrements and scans only them in the argmax; the CL local winners (with the position of
//     their lists) are exchanged through distributed shared memory, one hardware cluster barrier per step.  This
//     is what makes S > 65,535 fit (the per-sample state of 100,000 samples does not fit one SM), and it divides
//     the shared-memory atomics of heavy picks by CL.
template <int ESTRIDE, bool FAST, int CL, bool WIDE>
__global__ void __launch_bounds__(1024, 1) select_tail_kernel(SelParams p, TailCfg cfg, unsigned long long lists_total)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ Cand s_red[32];
    __shared__ unsigned long long s_sum[32];
    __shared__ uint4 s_xbest[2][16];               // CL > 1: the local winners of every CTA, double buffered by step parity
    __shared__ uint2 s_xlist[2][16];               //         ... and where their edge lists are (offset, length)
    __shared__ unsigned long long s_xsum[16];
    __shared__ unsigned int s_stage_n;
    constexpr bool AF = ESTRIDE == 2;
    constexpr int kInl = WIDE ? 2 : kInline;       // carriers inline in an entry
    constexpr int kPerChunk = WIDE ? 4 : 8;        // carriers per 16-byte pool chunk
    constexpr unsigned int kPad = WIDE ? 0xffffffffu : 0xffffu;
    int crank = 0;
    if (CL > 1) crank = (int)cg::this_cluster().block_rank();
    int xpar = 0;
    int want_single = 0;
    unsigned int *s_cnt = reinterpret_cast<unsigned int *>(smem);
    unsigned long long *s_lo = reinterpret_cast<unsigned long long *>(smem + cfg.off_lo);
    unsigned long long *s_hi = reinterpret_cast<unsigned long long *>(smem + cfg.off_hi);
    double *s_w = reinterpret_cast<double *>(smem + cfg.off_w);
    uint8_t *s_mask = smem + cfg.off_mask;
    unsigned int *s_loff = reinterpret_cast<unsigned int *>(smem + cfg.off_loff);     // CL == 1 only
    unsigned int *s_llen = reinterpret_cast<unsigned int *>(smem + cfg.off_llen);
    uint32_t *s_live = reinterpret_cast<uint32_t *>(smem + cfg.off_live);
    uint4 *s_stage = reinterpret_cast<uint4 *>(smem + cfg.off_stage);
    const uint4 *pool16 = reinterpret_cast<const uint4 *>(p.pool);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool has_w = p.weights != nullptr;
    const bool live_smem = cfg.live_words > 0;
    SelState *st = p.st;
    // samples this CTA owns: s = crank + CL * li, li = 0 .. n_own-1
    const int n_own = p.S > crank ? (p.S - crank + CL - 1) / CL : 0;
#define UT_OWNED(s) (CL == 1 || (int)((s) % CL) == crank)
#define UT_LI(s) (CL == 1 ? (s) : (s) / CL)

    for (int li = tid; li < n_own; li += blockDim.x) {
        const int i = crank + CL * li;
        s_cnt[li] = p.gain_cnt[i];
        s_mask[li] = p.mask[i];
        if (CL == 1) { s_loff[li] = p.list_off[i]; s_llen[li] = p.list_len[i]; }
        if (AF) { s_lo[li] = p.gain_lo[i]; s_hi[li] = p.gain_hi[i]; }
        if (has_w) s_w[li] = p.weights[i];
    }
    for (int i = tid; i < cfg.live_words; i += blockDim.x) s_live[i] = p.live[i];
    uint32_t *g_live = p.live;
    if (CL > 1 && !live_smem) {                       // every CTA of the cluster clears bits in its own copy
        g_live = cfg.live_priv + (size_t)crank * (size_t)p.colPitchW;
        for (long long i = tid; i < p.colPitchW; i += blockDim.x) g_live[i] = p.live[i];
    }
    long long step = st->step, tot = st->tot;
    const long long limit = st->limit;
    int stop = st->stop;
    int recompact = 0;
    int since_check = 0;
    if (tid == 0) s_stage_n = 0;
    __syncthreads();
    long long t_arg = 0, t_walk = 0, t_mark = clock64();
    unsigned int n_walked = 0, n_fresh = 0, n_inl = 0, n_pool = 0;       // work counters (profiling, p.dbg[5..7], [12])
#define UT_TICK(acc) do { const long long now__ = clock64(); acc += now__ - t_mark; t_mark = now__; } while (0)

    while (stop == 0 && step < limit) {
        // ---- every 64 steps: how many list entries are still live?  (sum of the gains)
        if (++since_check >= 64) {
            since_check = 0;
            unsigned long long acc = 0;
            for (int li = tid; li < n_own; li += blockDim.x) acc += s_mask[li] == 1 ? s_cnt[li] : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) s_sum[warp] = acc;
            __syncthreads();
            unsigned long long live_now = s_sum[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) live_now += __shfl_xor_sync(0xffffffffu, live_now, o);
            __syncthreads();
            if (CL > 1) {                                  // every CTA holds the gains of its own samples only
                cg::cluster_group cluster = cg::this_cluster();
                if (tid < CL) cluster.map_shared_rank(s_xsum, tid)[crank] = live_now;
                cluster.sync();
                live_now = lane < CL ? s_xsum[lane] : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) live_now += __shfl_xor_sync(0xffffffffu, live_now, o);
                cluster.sync();                            // s_xsum may be rewritten by the next check
            }
            if (live_now >= cfg.min_recompact && live_now * 2 <= lists_total && limit - step > 128) {
                recompact = 1;
                break;
            }
        }
        // ---- argmax over all samples (shared memory, np.argmax order)
        int best_idx, next_idx;             // the pick, and the best of the other warps' / CTAs' winners (a likely next pick)
        unsigned int best_cnt;
        double best_score;
        unsigned int best_off = 0, best_len = 0, next_off = 0, next_len = 0;
        if (FAST) {
            unsigned int bk = 0, bi = 0x7fffffffu;
            for (int li = tid; li < n_own; li += blockDim.x) {
                const unsigned int k = s_mask[li] == 1 ? s_cnt[li] : 0u;
                if (k > bk || bi == 0x7fffffffu) { bk = k; bi = (unsigned int)(crank + CL * li); }     // ascending: first index kept on ties
            }
            const uint2 w = warp_argmax_u32(bk, bi);
            if (lane == 0) *reinterpret_cast<uint2 *>(&s_red[warp]) = w;
            __syncthreads();
            uint2 mine = *reinterpret_cast<const uint2 *>(&s_red[lane]);
            uint2 b = warp_argmax_u32(mine.x, mine.y);
            uint2 olist = make_uint2(0u, 0u);
            if (CL > 1) {
                cg::cluster_group cluster = cg::this_cluster();
                if (tid < CL) {
                    uint2 ll = make_uint2(0u, 0u);
                    if (b.y != 0x7fffffffu) ll = make_uint2(__ldg(p.list_off + b.y), __ldg(p.list_len + b.y));
                    cluster.map_shared_rank(&s_xbest[xpar][0], tid)[crank] = make_uint4(b.x, b.y, 0u, 0u);
                    cluster.map_shared_rank(&s_xlist[xpar][0], tid)[crank] = ll;
                }
                cluster.sync();
                const uint4 o = lane < CL ? s_xbest[xpar][lane] : make_uint4(0u, 0x7fffffffu, 0u, 0u);
                olist = lane < CL ? s_xlist[xpar][lane] : make_uint2(0u, 0u);
                xpar ^= 1;
                mine = make_uint2(o.x, o.y);
                b = warp_argmax_u32(mine.x, mine.y);
            }
            const uint2 t2 = warp_argmax_u32(mine.y == b.y ? 0u : mine.x, mine.y == b.y ? 0x7fffffffu : mine.y);
            best_idx = (int)b.y;
            best_cnt = b.x;
            best_score = (double)b.x;
            next_idx = t2.x > 0u ? (int)t2.y : 0x7fffffff;
            if (CL > 1) {
                const int src_b = b.y != 0x7fffffffu ? (int)(b.y % CL) : 0, src_n = next_idx != 0x7fffffff ? next_idx % CL : 0;
                best_off = __shfl_sync(0xffffffffu, olist.x, src_b);
                best_len = __shfl_sync(0xffffffffu, olist.y, src_b);
                next_off = __shfl_sync(0xffffffffu, olist.x, src_n);
                next_len = __shfl_sync(0xffffffffu, olist.y, src_n);
            }
        } else {
            Cand b{0u, 0u, 0x7fffffff, 0u};
            for (int li = tid; li < n_own; li += blockDim.x) {
                const unsigned int c = s_cnt[li];
                double g = 0.0;
                if (s_mask[li] == 1) {
                    g = AF ? fixed_to_double(s_lo[li], s_hi[li], p.L, p.scale) : (double)c;
                    if (has_w) g *= s_w[li];
                }
                const unsigned long long k = score_key(g);
                Cand c2{(unsigned int)(k >> 32), (unsigned int)k, crank + CL * li, c};
                if (cand_better(c2, b)) b = c2;
            }
            b = warp_argmax(b);
            if (lane == 0) s_red[warp] = b;
            __syncthreads();
            Cand mine = s_red[lane];
            b = warp_argmax(mine);
            uint2 olist = make_uint2(0u, 0u);
            if (CL > 1) {
                cg::cluster_group cluster = cg::this_cluster();
                if (tid < CL) {
                    uint2 ll = make_uint2(0u, 0u);
                    if (b.idx != 0x7fffffff) ll = make_uint2(__ldg(p.list_off + b.idx), __ldg(p.list_len + b.idx));
                    cluster.map_shared_rank(&s_xbest[xpar][0], tid)[crank] = make_uint4(b.hi, b.lo, (unsigned int)b.idx, b.cnt);
                    cluster.map_shared_rank(&s_xlist[xpar][0], tid)[crank] = ll;
                }
                cluster.sync();
                const uint4 o = lane < CL ? s_xbest[xpar][lane] : make_uint4(0u, 0u, 0x7fffffffu, 0u);
                olist = lane < CL ? s_xlist[xpar][lane] : make_uint2(0u, 0u);
                xpar ^= 1;
                mine = Cand{o.x, o.y, (int)o.z, o.w};
                b = warp_argmax(mine);
            }
            Cand t2 = mine;
            if (mine.idx == b.idx) { t2.hi = 0u; t2.lo = 0u; t2.idx = 0x7fffffff; }
            t2 = warp_argmax(t2);
            best_idx = b.idx;
            best_cnt = b.cnt;
            best_score = key_score(b.hi, b.lo);
            next_idx = (t2.idx != 0x7fffffff && key_score(t2.hi, t2.lo) > 0.0) ? t2.idx : 0x7fffffff;
            if (CL > 1) {
                const int src_b = best_idx != 0x7fffffff ? best_idx % CL : 0, src_n = next_idx != 0x7fffffff ? next_idx % CL : 0;
                best_off = __shfl_sync(0xffffffffu, olist.x, src_b);
                best_len = __shfl_sync(0xffffffffu, olist.y, src_b);
                next_off = __shfl_sync(0xffffffffu, olist.x, src_n);
                next_len = __shfl_sync(0xffffffffu, olist.y, src_n);
            }
        }
        if (p.S == 0 || best_score == 0.0) {              // utmos/select.py:51-52
            stop = UTMOS_STOP_ZERO;
            break;
        }
        if (CL > 1 && best_cnt < cfg.single_rows) {       // light picks: one CTA without a cluster barrier is faster
            want_single = 1;
            break;
        }
        if (CL == 1) {
            best_off = s_loff[best_idx];
            best_len = s_llen[best_idx];
            if (next_idx != 0x7fffffff) { next_off = s_loff[next_idx]; next_len = s_llen[next_idx]; }
        }
        if (tid == 0) {
            if (crank == 0) {
                p.out_idx[step] = best_idx;
                p.out_new[step] = best_cnt;
                p.out_score[step] = best_score;
                if (p.dbg_time) p.out_time[step] = global_timer_ns();
            }
            if (UT_OWNED(best_idx)) s_mask[UT_LI(best_idx)] = 0;      // utmos/select.py:100
        }
        step += 1;
        tot += best_cnt;
        if (tot >= p.V) {                                 // utmos/select.py:110-112
            stop = UTMOS_STOP_ALL;
            break;
        }
        UT_TICK(t_arg);
        // ---- stream the winner's list; a live bit that we clear marks a newly covered row
        const uint4 *lst = p.lists + (size_t)best_off * ESTRIDE;
        const int len = (int)best_len;
        if (next_idx != 0x7fffffff) {                     // warm L2 with the runner-up's list
            const uint4 *l2 = p.lists + (size_t)next_off * ESTRIDE;
            const int lines = ((int)next_len * ESTRIDE + 7) >> 3;    // 128-byte lines
            for (int i = tid; i < lines; i += blockDim.x) prefetch_l2(l2 + (size_t)i * 8);
        }
        const int sub = lane & 7, slot = lane >> 3;
        int staged_any = 0;
        // subtract one pooled chunk (16 bytes of carriers) from the gains this CTA owns
        auto retire_chunk = [&](const uint4 &v, unsigned long long gl, unsigned long long gh) {
            const unsigned int ww[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < kPerChunk; ++q) {
                const unsigned int cs = WIDE ? ww[q] : ((q & 1) ? ww[q >> 1] >> 16 : ww[q >> 1] & 0xffffu);
                if (cs != kPad && UT_OWNED(cs)) {
                    n_pool += 1;
                    atomicAdd(s_cnt + UT_LI(cs), 0xffffffffu);
                    if (AF) { smem_add64(s_lo + UT_LI(cs), gl); smem_add64(s_hi + UT_LI(cs), gh); }
                }
            }
        };
        // retire the staged carrier lists: one chunk per thread and turn
        auto retire_staged = [&]() {
            const unsigned int staged_n = min(s_stage_n, cfg.stage_cap);
            for (unsigned int c0 = tid; c0 < staged_n; c0 += blockDim.x) {
                const uint4 v = s_stage[(size_t)c0 * ESTRIDE];
                unsigned long long gl = 0, gh = 0;
                if (AF) {
                    const uint4 qv = s_stage[(size_t)c0 * ESTRIDE + 1];
                    gl = ((unsigned long long)qv.y << 32) | qv.x;
                    gh = ((unsigned long long)qv.w << 32) | qv.z;
                }
                retire_chunk(v, gl, gh);
            }
        };
        // a pick that covers very many rows is walked in smaller batches (one entry per thread instead of four) and
        // the staging area is drained after every batch, so that it does not overflow into the slow direct path
        const int ept = best_cnt > cfg.stage_cap / 2u ? 1 : 4;
        const int bstride = ept * (int)blockDim.x;
        for (int base = 0; base < len; base += bstride) {
            uint4 e[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = base + u * (int)blockDim.x + tid;
                e[u] = (u < ept && i < len) ? __ldg(lst + (size_t)i * ESTRIDE) : make_uint4(0xffffffffu, 0u, 0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (u >= ept || base + u * (int)blockDim.x + warp * 32 >= len) break;       // warp-uniform
                const int i = base + u * (int)blockDim.x + tid;
                const unsigned int r = e[u].x;
                bool fresh = false;
                if (r != 0xffffffffu) {
                    // a row appears once in a list, so nobody else clears this bit during the walk: test with a plain
                    // load (shared-memory atomics cost ~2 cycles per lane) and clear only the bits that are set
                    const uint32_t bit = 1u << (r & 31);
                    if (live_smem) {
                        uint32_t *lw = s_live + (r >> 5);
                        fresh = (*reinterpret_cast<volatile uint32_t *>(lw) & bit) != 0;
                        if (fresh) atomicAnd(lw, ~bit);
                    } else {
                        // live mask in global memory (more rows than one SM holds, e.g. the merged rows of several GPUs):
                        // the old value is not needed, so clear the bit with a fire-and-forget RED instead of an atomic
                        // that waits for its answer (ncu r2: 22 % of the stall samples of this kernel sat on that wait)
                        uint32_t *lw = g_live + (r >> 5);
                        uint32_t cur;
                        asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(cur) : "l"(lw) : "memory");
                        fresh = (cur & bit) != 0;
                        if (fresh) asm volatile("red.global.and.b32 [%0], %1;" ::"l"(lw), "r"(~bit) : "memory");
                    }
                }
                // narrow entry: {row, n | c0 << 16, c1 | c2 << 16, c3 | c4 << 16}; wide entry: {row, n, c0, c1};
                // pooled (n == all ones): {row, pooled, first pool element, carriers}
                const unsigned int n = WIDE ? e[u].y : (e[u].y & 0xffffu);
                const bool pooled = WIDE ? n == 0xffffffffu : n == kPooled;
                unsigned long long nl = 0, nh = 0;
                if (AF && fresh) {
                    const uint4 qv = __ldg(lst + (size_t)i * ESTRIDE + 1);
                    nl = 0ull - (((unsigned long long)qv.y << 32) | qv.x);
                    nh = 0ull - (((unsigned long long)qv.w << 32) | qv.z);
                }
                n_walked += r != 0xffffffffu;
                n_fresh += fresh;
                if (fresh && !pooled) {
                    n_inl += n;
                    unsigned int c[kInline];
                    if (WIDE) { c[0] = e[u].z; c[1] = e[u].w; c[2] = c[3] = c[4] = 0; }
                    else { c[0] = e[u].y >> 16; c[1] = e[u].z & 0xffffu; c[2] = e[u].z >> 16; c[3] = e[u].w & 0xffffu; c[4] = e[u].w >> 16; }
#pragma unroll
                    for (int j = 0; j < kInl; ++j) {
                        if (j < (int)n && UT_OWNED(c[j])) {
                            atomicAdd(s_cnt + UT_LI(c[j]), 0xffffffffu);
                            if (AF) { smem_add64(s_lo + UT_LI(c[j]), nl); smem_add64(s_hi + UT_LI(c[j]), nh); }
                        }
                    }
                }
                // rows with many carriers keep their carrier list (padded to whole 16-byte chunks, 16-byte aligned) in
                // the pool.  The lists of all the rows this pick newly covers are copied into shared memory with
                // cp.async -- every copy in flight at once, one global round trip whatever the number of rows -- and
                // retired from there after the walk, flat over all threads.
                const bool big = fresh && pooled;
                unsigned int m = 0;
                if (__any_sync(0xffffffffu, big)) {
                    staged_any = 1;
                    // reserve chunks of the staging area (one shared-memory atomic per warp instruction: the lanes'
                    // same-address adds are serialised by the hardware, which is cheaper than a warp scan)
                    const unsigned int n8 = big ? (e[u].w + kPerChunk - 1) / kPerChunk : 0u;
                    const unsigned int my0 = big ? atomicAdd(&s_stage_n, n8) : 0u;
                    const bool staged = big && my0 + n8 <= cfg.stage_cap;
                    if (staged) {
                        const uint4 *src = pool16 + e[u].z / kPerChunk;
                        for (unsigned int k = 0; k < n8; ++k) {
                            uint4 *dst = s_stage + (size_t)(my0 + k) * ESTRIDE;
                            const unsigned int sa = (unsigned int)__cvta_generic_to_shared(dst);
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src + k) : "memory");
                            if (AF) dst[1] = make_uint4((unsigned int)nl, (unsigned int)(nl >> 32), (unsigned int)nh, (unsigned int)(nh >> 32));
                        }
                    } else if (big) {
                        // did not fit: blank the part of the reservation that lies inside the staging area
                        for (unsigned int k = my0; k < my0 + n8 && k < cfg.stage_cap; ++k)
                            s_stage[(size_t)k * ESTRIDE] = make_uint4(~0u, ~0u, ~0u, ~0u);
                    }
                    m = __ballot_sync(0xffffffffu, big && !staged);
                }
                // overflow (a pick that covers more carriers than the staging area holds): the warp retires the row from
                // global memory, four rows at a time, 8 lanes x 128-bit loads per row
                while (m) {
                    int src = -1;
#pragma unroll
                    for (int s4 = 0; s4 < 4; ++s4) {
                        if (m) {
                            const int bpos = __ffs(m) - 1;
                            m &= m - 1;
                            if (slot == s4) src = bpos;
                        }
                    }
                    const int from = src < 0 ? 0 : src;
                    const unsigned int pbase = __shfl_sync(0xffffffffu, e[u].z, from);
                    const unsigned int cnt_from = __shfl_sync(0xffffffffu, e[u].w, from);   // every lane takes part
                    const unsigned int cnt = src < 0 ? 0u : cnt_from;
                    unsigned long long gl = 0, gh = 0;
                    if (AF) { gl = __shfl_sync(0xffffffffu, nl, from); gh = __shfl_sync(0xffffffffu, nh, from); }
                    const uint4 *pl = pool16 + pbase / kPerChunk;
                    const unsigned int n8 = (cnt + kPerChunk - 1) / kPerChunk;
                    for (unsigned int k0 = sub; k0 < n8; k0 += 16) {
                        const uint4 v0 = __ldg(pl + k0);
                        const uint4 v1 = k0 + 8 < n8 ? __ldg(pl + k0 + 8) : make_uint4(~0u, ~0u, ~0u, ~0u);
                        retire_chunk(v0, gl, gh);
                        retire_chunk(v1, gl, gh);
                    }
                }
            }
            if (base + bstride < len) {                   // more batches follow (block-uniform): drain the staging area
                asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
                if (__syncthreads_or(staged_any)) {
                    retire_staged();
                    __syncthreads();
                    if (tid == 0) s_stage_n = 0;
                    __syncthreads();
                }
                staged_any = 0;
            }
        }
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
        if (__syncthreads_or(staged_any)) {               // nothing staged (late tail): this barrier already ends the step
            retire_staged();
            __syncthreads();
        }
        // every live row of the pick is covered now: its own gain is zero (keeps sum(gains) == live list entries)
        if (tid == 0) {
            s_stage_n = 0;                                // the next walk starts after the argmax barrier
            if (UT_OWNED(best_idx)) {
                s_cnt[UT_LI(best_idx)] = 0;
                if (AF) { s_lo[UT_LI(best_idx)] = 0; s_hi[UT_LI(best_idx)] = 0; }
            }
        }
        UT_TICK(t_walk);
    }
#undef UT_TICK
    if (tid == 0 && p.dbg && crank == 0) { p.dbg[8] += t_arg; p.dbg[9] += t_walk; p.dbg[11] += 1; }
    if (p.dbg && crank == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            n_walked += __shfl_xor_sync(0xffffffffu, n_walked, o);
            n_fresh += __shfl_xor_sync(0xffffffffu, n_fresh, o);
            n_inl += __shfl_xor_sync(0xffffffffu, n_inl, o);
            n_pool += __shfl_xor_sync(0xffffffffu, n_pool, o);
        }
        if (lane == 0) {
            atomicAdd(reinterpret_cast<unsigned long long *>(p.dbg + 5), (unsigned long long)n_walked);
            atomicAdd(reinterpret_cast<unsigned long long *>(p.dbg + 6), (unsigned long long)n_fresh);
            atomicAdd(reinterpret_cast<unsigned long long *>(p.dbg + 7), (unsigned long long)n_inl);
            atomicAdd(reinterpret_cast<unsigned long long *>(p.dbg + 12), (unsigned long long)n_pool);
        }
    }

    __syncthreads();
    for (int li = tid; li < n_own; li += blockDim.x) {       // every CTA: the samples it owns
        const int i = crank + CL * li;
        p.gain_cnt[i] = s_cnt[li];
        p.mask[i] = s_mask[li];
        if (AF) { p.gain_lo[i] = s_lo[li]; p.gain_hi[i] = s_hi[li]; }
    }
    if (crank == 0) {
        for (int i = tid; i < cfg.live_words; i += blockDim.x) p.live[i] = s_live[i];
        if (CL > 1 && !live_smem)
            for (long long i = tid; i < p.colPitchW; i += blockDim.x) p.live[i] = g_live[i];
        if (tid == 0) {
            st->step = step;
            st->tot = tot;
            st->stop = stop;
            st->winner = -1;
            st->regain = 0;
            st->recompact = recompact;
            if (want_single) st->tail_single |= 1u;
        }
    }
#undef UT_OWNED
#undef UT_LI
    if (CL > 1) cg::this_cluster().sync();           // nobody leaves while a peer may still write into its shared memory
}

// ------------------------------------------------------------------------------------------------
// List-driven tail, ENTRY-DIVIDED over a cluster: the entries of the pick split over 16 CTAs, the per-sample state sliced
// over their shared memories, decrements sent to the owner's slice with remote shared-memory atomics.
//
// select_tail_kernel applies every decrement of a step with the shared-memory atomics of ONE SM (2 cycles per lane), and its
// owner-computes cluster flavour makes every CTA walk ALL the entries of the pick.  That is the right trade while a pick
// covers a few hundred rows and the wrong one when it covers thousands: the merged lists of N GPUs hold N times the rows
// per pick (8 GPUs: 12 M decrements = 18 ms of a 20 ms tail), and so does one GPU with a large matrix.  Here
//   * CTA r keeps the gains / mask / list positions of the samples s with s % CL == r in its shared memory (loaded from
//     global at entry, written back at exit) and scans only those for its local best;
//   * the CL local winners (with their list position) meet through distributed shared memory, one cluster barrier;
//   * the pick's entries are dealt out warp by warp over all CL x 32 warps of the cluster; the live bit is cleared with a
//     global atomic that returns the old word (that decides "newly covered"); each carrier's decrement is a remote
//     shared-memory atomic on the owner's slice (64-bit limbs as two 32-bit atomics with carry, like smem_add64); pooled
//     carrier lists are queued in shared memory and retired a warp per row;
//   * a second cluster barrier (release/acquire over shared::cluster) ends the step: no fence, no L2 round trip for gains.
// Same lists, same pool, same recompaction protocol, same pick order (np.argmax, utmos/select.py:48) as select_tail_kernel.
// ------------------------------------------------------------------------------------------------
// shared::cluster address of `local` (a pointer into this CTA's shared memory) in the CTA `rank` of the cluster
__device__ __forceinline__ uint32_t dsmem_addr(const void *local, unsigned int rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"((uint32_t)__cvta_generic_to_shared(local)), "r"(rank));
    return r;
}
__device__ __forceinline__ void dsmem_red_add(uint32_t addr, unsigned int v)
{
    asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// 16 bytes into a peer's shared memory, counted on the PEER's mbarrier when they have landed (st.async): the winners'
// exchange needs no cluster barrier (a MEMBAR.ALL.GPU + arrive/wait + L1 invalidate, ~1,100 cycles), only this round trip
__device__ __forceinline__ void dsmem_store16_tx(uint32_t addr, uint4 v, uint32_t bar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(bar) : "memory");
}
__device__ __forceinline__ void lc_mbar_init(unsigned long long *bar, unsigned int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void lc_mbar_expect_tx(unsigned long long *bar, unsigned int bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void lc_mbar_wait(unsigned long long *bar, unsigned int parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "LC_WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni LC_WAIT_DONE;\n\t"
        "bra.uni LC_WAIT_LOOP;\n\t"
        "LC_WAIT_DONE:\n\t"
        "}" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}

// 64-bit add on a slice of the cluster's shared memory.  SASS: one ATOM.E.ADD.64 that the memory system performs when the
// address is a peer's (15 of 16 cases; nothing comes back), and the ATOMS.CAST.SPIN compare-and-swap loop when it is this
// CTA's own.  (Two 32-bit atomics with carry, as in smem_add64, need the old low word back: a remote round trip per limb.)
__device__ __forceinline__ void dsmem_add64(uint32_t addr, unsigned long long v)
{
    asm volatile("red.relaxed.cluster.shared::cluster.add.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}

struct ListClusterCfg {
    int n_own_max;       // samples per CTA slice
    int off_lo, off_hi, off_mask, off_loff, off_llen;   // byte offsets into dynamic shared memory (counts at 0)
    int stage_lists;     // list positions staged in shared memory (else read through the read-only path)
    unsigned int min_recompact;
    int off_rows, off_vals;      // REFT: sort buffer of one candidate's live rows / their AF values
    unsigned int row_cap;        // REFT: rows the buffers hold (a power of two; 0 = no room, REFT not available)
};

// FAST: count mode without weights -> candidates are (gain count of a selectable sample, index): two REDUX per reduction
template <bool FAST>
__device__ __forceinline__ Cand lc_warp_best(Cand c)
{
    if (FAST) {
        const uint2 w = warp_argmax_u32(c.lo, (unsigned int)c.idx);
        return Cand{0u, w.x, (int)w.y, w.x};
    }
    return warp_argmax(c);
}

// REFT (--af with UTMOS_F_REF_TIES): the reference's tie order inside the tail.  The exact fixed-point scores decide a step
// unless two or more samples lie within 2^-30 (relative) of the best one; then every such candidate's score is replayed the
// way the reference computes it -- float64 AF of its uncovered rows added one after the other in ascending row order
// (utmos/select.py:37-40), times the weight (:47) -- and the first maximum of the replayed values wins (:48), exactly as in
// argmax_step_kernel.  Here the candidate's uncovered rows come from its edge list (filtered by the live mask), are sorted
// in shared memory (bitonic, one CTA per candidate, the candidates of different CTAs in parallel) and summed by one thread.
// A candidate with more live rows than the sort buffer holds ends the launch with st->tie_step set: the host runs that one
// step with the per-step kernels and comes back.
template <int ESTRIDE, bool FAST, int CL, bool WIDE, bool REFT>
__global__ void __launch_bounds__(1024, 1) select_listcluster_kernel(SelParams p, ListClusterCfg cfg, unsigned long long lists_total,
                                                                     unsigned int light_rows)
{
    constexpr bool AF = ESTRIDE == 2;
    constexpr int kInl = WIDE ? 2 : kInline;
    constexpr unsigned int kPad = WIDE ? 0xffffffffu : 0xffffu;
    constexpr int kQueue = 2048;
    extern __shared__ __align__(16) unsigned char lc_smem[];
    __shared__ Cand s_red[32];
    __shared__ unsigned long long s_sum[32];
    __shared__ uint4 s_xbest[2][16];
    __shared__ uint4 s_xlist[2][16];                 // {sum of owned gains lo, hi, list offset, list length}
    __shared__ uint4 s_q[kQueue];                    // pooled rows this CTA found: {pool base, carriers, entry index, 0}
    __shared__ unsigned int s_qn[2];                 // pooled rows queued in this step / zeroed for the next one
    __shared__ __align__(8) unsigned long long s_xbar;   // counts the 2 x CL 16-byte slots of one exchange
    __shared__ __align__(8) unsigned long long s_xbar2, s_xbar3;   // REFT: candidate counts / replayed winners
    __shared__ uint4 s_xcnt[2][16];                  // REFT: {candidates within the tie window in that CTA's slice, 0, 0, 0}
    __shared__ uint4 s_xrep[2][16], s_xrepl[2][16];  // REFT: replayed local winner {key hi, key lo, sample, count}, {list offset, length, 0, 0}
    __shared__ unsigned int s_ncand, s_nrow;
    unsigned int *s_cnt = reinterpret_cast<unsigned int *>(lc_smem);
    unsigned long long *s_lo = reinterpret_cast<unsigned long long *>(lc_smem + cfg.off_lo);
    unsigned long long *s_hi = reinterpret_cast<unsigned long long *>(lc_smem + cfg.off_hi);
    unsigned char *s_mask = lc_smem + cfg.off_mask;
    unsigned int *s_loff = reinterpret_cast<unsigned int *>(lc_smem + cfg.off_loff);
    unsigned int *s_llen = reinterpret_cast<unsigned int *>(lc_smem + cfg.off_llen);
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool has_w = p.weights != nullptr;
    SelState *st = p.st;
    const int n_own = p.S > crank ? (p.S - crank + CL - 1) / CL : 0;
    long long step = st->step, tot = st->tot;
    const long long limit = st->limit;
    int stop = st->stop, recompact = 0, want_light = 0, xpar = 0, since_check = 0, ph2 = 0, ph3 = 0, tie_fallback = 0;
    unsigned int *s_rows = reinterpret_cast<unsigned int *>(lc_smem + cfg.off_rows);     // REFT: live rows of one candidate
    double *s_vals = reinterpret_cast<double *>(lc_smem + cfg.off_vals);                 // REFT: their AF values, row order
    int *s_cand = reinterpret_cast<int *>(s_q);                                          // REFT: slots of this CTA's candidates
    const unsigned int *pool32 = reinterpret_cast<const unsigned int *>(p.pool);
    for (int li = tid; li < n_own; li += blockDim.x) {
        const int s = crank + CL * li;
        s_cnt[li] = p.gain_cnt[s];
        if (AF) { s_lo[li] = p.gain_lo[s]; s_hi[li] = p.gain_hi[s]; }
        s_mask[li] = p.mask[s];
        if (cfg.stage_lists) { s_loff[li] = p.list_off[s]; s_llen[li] = p.list_len[s]; }
    }
    if (tid == 0) {
        lc_mbar_init(&s_xbar, 1u);
        if (REFT) { lc_mbar_init(&s_xbar2, 1u); lc_mbar_init(&s_xbar3, 1u); }
        s_qn[0] = s_qn[1] = 0u;
    }
    __syncthreads();
    cluster.sync();
    // the slice of sample c: CTA c % CL, slot c / CL
    auto retire = [&](unsigned int c, unsigned long long nl, unsigned long long nh) {
        const unsigned int owner = c % CL, slot = c / CL;
        dsmem_red_add(dsmem_addr(s_cnt + slot, owner), 0xffffffffu);
        if (AF) {
            dsmem_add64(dsmem_addr(s_lo + slot, owner), nl);
            dsmem_add64(dsmem_addr(s_hi + slot, owner), nh);
        }
    };

    // integer keys: one warp scans a slice of up to 512 samples alone (16 per lane) -- no block barrier, no second reduction
    // level.  Float64 scores (a 128-bit fixed-point conversion per sample) want every warp: C3 tail 5.75 ms against 6.86 ms.
    const bool one_warp = n_own <= (FAST ? 512 : 32);
    while (stop == 0 && step < limit) {
        // ---- argmax over the samples this CTA owns; the sum of their gains only on the steps that look at it
        const bool check = since_check + 1 >= 64;
        const int qp = xpar;
        Cand b{0u, 0u, 0x7fffffff, 0u};
        unsigned long long acc = 0;
        if (tid == 0) {
            s_qn[qp ^ 1] = 0;                         // next step's queue: last read before the barrier that ended the previous step
            lc_mbar_expect_tx(&s_xbar, (unsigned int)CL * 32u);
        }
        if (!one_warp || warp == 0) {
            for (int li = one_warp ? lane : tid; li < n_own; li += one_warp ? 32 : (int)blockDim.x) {
                const int s = crank + CL * li;
                const unsigned int c = s_cnt[li];
                const bool ok = s_mask[li] == 1;
                Cand c2;
                if (FAST) {
                    c2 = Cand{0u, ok ? c : 0u, s, ok ? c : 0u};
                } else {
                    double g = 0.0;
                    if (ok) {
                        g = AF ? fixed_to_double(s_lo[li], s_hi[li], p.L, p.scale) : (double)c;
                        if (has_w) g *= __ldg(p.weights + s);
                    }
                    const unsigned long long kk = score_key(g);
                    c2 = Cand{(unsigned int)(kk >> 32), (unsigned int)kk, s, c};
                }
                if (ok) acc += c;
                if (cand_better(c2, b)) b = c2;
            }
            b = lc_warp_best<FAST>(b);
            if (check) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            }
        }
        if (!one_warp) {
            if (lane == 0) { s_red[warp] = b; s_sum[warp] = acc; }
            __syncthreads();
            if (warp == 0) {
                b = lc_warp_best<FAST>(s_red[lane]);
                if (check) {
                    acc = s_sum[lane];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                }
            }
        }
        if (warp == 0 && lane < CL) {
            unsigned int off = 0, len = 0;
            if (b.idx != 0x7fffffff) {
                const int li = (b.idx - crank) / CL;
                off = cfg.stage_lists ? s_loff[li] : __ldg(p.list_off + b.idx);
                len = cfg.stage_lists ? s_llen[li] : __ldg(p.list_len + b.idx);
            }
            const uint32_t bar = dsmem_addr(&s_xbar, (unsigned int)lane);
            dsmem_store16_tx(dsmem_addr(&s_xbest[xpar][crank], (unsigned int)lane), make_uint4(b.hi, b.lo, (unsigned int)b.idx, b.cnt), bar);
            dsmem_store16_tx(dsmem_addr(&s_xlist[xpar][crank], (unsigned int)lane),
                             make_uint4((unsigned int)acc, (unsigned int)(acc >> 32), off, len), bar);
        }
        lc_mbar_wait(&s_xbar, (unsigned int)xpar);   // exchange number n completes phase n of the mbarrier: parity = xpar
        unsigned int off, len;
        {
            const uint4 o = lane < CL ? s_xbest[xpar][lane] : make_uint4(0u, 0u, 0x7fffffffu, 0u);
            const uint4 l = lane < CL ? s_xlist[xpar][lane] : make_uint4(0u, 0u, 0u, 0u);
            b = lc_warp_best<FAST>(Cand{o.x, o.y, (int)o.z, o.w});
            if (check) {
                acc = ((unsigned long long)l.y << 32) | l.x;
#pragma unroll
                for (int o2 = 16; o2 > 0; o2 >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o2);
            }
            const int src = b.idx == 0x7fffffff ? 0 : b.idx % CL;             // the winner's CTA holds its list position
            off = __shfl_sync(0xffffffffu, l.z, src);
            len = __shfl_sync(0xffffffffu, l.w, src);
            xpar ^= 1;
        }
        if (REFT && b.idx != 0x7fffffff && key_score(b.hi, b.lo) > 0.0) {
            // ---- how many samples lie within the tie window of the exact winner?  (one: it stands)
            const double thr = key_score(b.hi, b.lo) * (1.0 - 9.313225746154785e-10);         // 2^-30
            auto in_window = [&](int li) {
                if (s_mask[li] != 1) return false;
                double g = fixed_to_double(s_lo[li], s_hi[li], p.L, p.scale);
                if (has_w) g *= __ldg(p.weights + crank + CL * li);
                return g >= thr;
            };
            int mine = 0;
            for (int li = tid; li < n_own; li += blockDim.x) mine += in_window(li) ? 1 : 0;
            const int holders = __syncthreads_count(mine > 0);
            const int crowded = __syncthreads_or(mine > 1);
            if (tid == 0) lc_mbar_expect_tx(&s_xbar2, (unsigned int)CL * 16u);
            if (warp == 0 && lane < CL)
                dsmem_store16_tx(dsmem_addr(&s_xcnt[ph2][crank], (unsigned int)lane),
                                 make_uint4((unsigned int)holders + (crowded ? 1u : 0u), 0u, 0u, 0u), dsmem_addr(&s_xbar2, (unsigned int)lane));
            lc_mbar_wait(&s_xbar2, (unsigned int)ph2);
            unsigned int total = lane < CL ? s_xcnt[ph2][lane].x : 0u;
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o2);
            ph2 ^= 1;
            if (total >= 2u) {
                // ---- replay: this CTA's candidates one after the other, the CTAs in parallel
                if (tid == 0) { s_ncand = 0; lc_mbar_expect_tx(&s_xbar3, (unsigned int)CL * 32u); }
                __syncthreads();
                for (int li = tid; li < n_own; li += blockDim.x)
                    if (in_window(li)) s_cand[atomicAdd(&s_ncand, 1u)] = li;
                __syncthreads();
                const int ncand = (int)s_ncand;
                Cand rb{0u, 0u, 0x7fffffff, 0u};
                unsigned int roff = 0, rlen = 0;
                int overflow = 0;
                for (int ci = 0; ci < ncand; ++ci) {
                    const int li = s_cand[ci], t = crank + CL * li;
                    const unsigned int toff = cfg.stage_lists ? s_loff[li] : __ldg(p.list_off + t);
                    const unsigned int tlen = cfg.stage_lists ? s_llen[li] : __ldg(p.list_len + t);
                    const uint4 *tl = p.lists + (size_t)toff * ESTRIDE;
                    if (tid == 0) s_nrow = 0;
                    __syncthreads();
                    for (unsigned int i = tid; i < tlen; i += blockDim.x) {
                        const unsigned int r = __ldg(tl + (size_t)i * ESTRIDE).x;
                        if ((__ldcg(p.live + (r >> 5)) >> (r & 31)) & 1u) {
                            const unsigned int slot = atomicAdd(&s_nrow, 1u);
                            if (slot < cfg.row_cap) s_rows[slot] = r;
                        }
                    }
                    __syncthreads();
                    const unsigned int n = s_nrow;
                    if (n > cfg.row_cap) { overflow = 1; continue; }
                    unsigned int N = 1;
                    while (N < n) N <<= 1;
                    for (unsigned int i = n + tid; i < N; i += blockDim.x) s_rows[i] = 0xffffffffu;
                    __syncthreads();
                    for (unsigned int k2 = 2; k2 <= N; k2 <<= 1) {
                        for (unsigned int j = k2 >> 1; j > 0; j >>= 1) {
                            for (unsigned int i = tid; i < N; i += blockDim.x) {
                                const unsigned int x = i ^ j;
                                if (x > i) {
                                    const unsigned int va = s_rows[i], vb = s_rows[x];
                                    if ((va > vb) == ((i & k2) == 0)) { s_rows[i] = vb; s_rows[x] = va; }
                                }
                            }
                            __syncthreads();
                        }
                    }
                    for (unsigned int i = tid; i < n; i += blockDim.x) {
                        double v = __ldg(p.af_vals + s_rows[i]);
                        if (p.af_f32) v = (double)(float)v;                   // hdf5 flavour: float32 GT*AF rows (utmos/select.py:218-223)
                        s_vals[i] = v;
                    }
                    __syncthreads();
                    if (tid == 0) {
                        double acc2 = 0.0;
                        for (unsigned int i = 0; i < n; ++i) acc2 += s_vals[i];       // strictly sequential, ascending rows
                        if (has_w) acc2 *= __ldg(p.weights + t);
                        const unsigned long long kk = score_key(acc2);
                        const Cand c2{(unsigned int)(kk >> 32), (unsigned int)kk, t, s_cnt[li]};
                        if (cand_better(c2, rb)) { rb = c2; roff = toff; rlen = tlen; }
                    }
                    __syncthreads();
                }
                if (warp == 0) {
                    rb.hi = __shfl_sync(0xffffffffu, rb.hi, 0);
                    rb.lo = __shfl_sync(0xffffffffu, rb.lo, 0);
                    rb.idx = __shfl_sync(0xffffffffu, rb.idx, 0);
                    rb.cnt = __shfl_sync(0xffffffffu, rb.cnt, 0);
                    roff = __shfl_sync(0xffffffffu, roff, 0);
                    rlen = __shfl_sync(0xffffffffu, rlen, 0);
                    if (lane < CL) {
                        const uint32_t bar = dsmem_addr(&s_xbar3, (unsigned int)lane);
                        dsmem_store16_tx(dsmem_addr(&s_xrep[ph3][crank], (unsigned int)lane),
                                         make_uint4(rb.hi, rb.lo, overflow ? 0x7ffffffeu : (unsigned int)rb.idx, rb.cnt), bar);
                        dsmem_store16_tx(dsmem_addr(&s_xrepl[ph3][crank], (unsigned int)lane), make_uint4(roff, rlen, 0u, 0u), bar);
                    }
                }
                lc_mbar_wait(&s_xbar3, (unsigned int)ph3);
                {
                    const uint4 o = lane < CL ? s_xrep[ph3][lane] : make_uint4(0u, 0u, 0x7fffffffu, 0u);
                    const uint4 l = lane < CL ? s_xrepl[ph3][lane] : make_uint4(0u, 0u, 0u, 0u);
                    ph3 ^= 1;
                    if (__any_sync(0xffffffffu, o.z == 0x7ffffffeu)) { tie_fallback = 1; break; }
                    b = warp_argmax(Cand{o.x, o.y, (int)o.z, o.w});
                    const int src = b.idx == 0x7fffffff ? 0 : b.idx % CL;
                    off = __shfl_sync(0xffffffffu, l.x, src);
                    len = __shfl_sync(0xffffffffu, l.y, src);
                }
            }
        }
        const int best_idx = b.idx;
        const unsigned int best_cnt = b.cnt;
        const double best_score = FAST ? (double)b.lo : key_score(b.hi, b.lo);
        if (p.S == 0 || best_idx == 0x7fffffff || best_score == 0.0) {        // utmos/select.py:51-52
            stop = UTMOS_STOP_ZERO;
            break;
        }
        if (++since_check >= 64) {
            since_check = 0;
            if (acc >= cfg.min_recompact && acc * 2 <= lists_total && limit - step > 128) { recompact = 1; break; }
        }
        if (light_rows && best_cnt < light_rows) { want_light = 1; break; }   // light picks: one SM from shared memory is faster
        if (tid == 0) {
            if (crank == 0) {
                p.out_idx[step] = best_idx;
                p.out_new[step] = best_cnt;
                p.out_score[step] = best_score;
                if (p.dbg_time) p.out_time[step] = global_timer_ns();
            }
            if (best_idx % CL == crank) {
                // every live row of the pick is covered by the end of this step and nobody decrements the pick itself
                const int li = best_idx / CL;
                s_cnt[li] = 0u;
                if (AF) { s_lo[li] = 0ull; s_hi[li] = 0ull; }
                s_mask[li] = 0;                                               // utmos/select.py:100
            }
        }
        step += 1;
        tot += best_cnt;
        if (tot >= p.V) {                                                     // utmos/select.py:110-112
            stop = UTMOS_STOP_ALL;
            break;
        }
        // ---- walk: the pick's entries dealt out warp by warp over the whole cluster
        const uint4 *lst = p.lists + (size_t)off * ESTRIDE;
        for (unsigned int i = ((unsigned int)warp * CL + (unsigned int)crank) * 32u + (unsigned int)lane; i < len;
             i += (unsigned int)CL * 1024u) {
            const uint4 e = __ldg(lst + (size_t)i * ESTRIDE);
            const unsigned int r = e.x;
            const uint32_t bit = 1u << (r & 31);
            const uint32_t old = atomicAnd(p.live + (r >> 5), ~bit);
            if (!(old & bit)) continue;                                       // covered earlier
            const unsigned int n = WIDE ? e.y : (e.y & 0xffffu);
            const bool pooled = WIDE ? n == 0xffffffffu : n == kPooled;
            if (!pooled) {
                unsigned int c[kInline];
                if (WIDE) { c[0] = e.z; c[1] = e.w; c[2] = c[3] = c[4] = 0; }
                else { c[0] = e.y >> 16; c[1] = e.z & 0xffffu; c[2] = e.z >> 16; c[3] = e.w & 0xffffu; c[4] = e.w >> 16; }
                unsigned long long nl = 0, nh = 0;
                if (AF && n) {
                    const uint4 qv = __ldg(lst + (size_t)i * ESTRIDE + 1);
                    nl = 0ull - (((unsigned long long)qv.y << 32) | qv.x);
                    nh = 0ull - (((unsigned long long)qv.w << 32) | qv.z);
                }
#pragma unroll
                for (int j = 0; j < kInl; ++j)
                    if (j < (int)n) retire(c[j], nl, nh);
            } else {
                const unsigned int slot = atomicAdd(&s_qn[qp], 1u);
                if (slot < (unsigned int)kQueue) {
                    s_q[slot] = make_uint4(e.z, e.w, i, 0u);
                } else {                                                      // queue full: this thread retires the row itself
                    unsigned long long nl = 0, nh = 0;
                    if (AF) {
                        const uint4 qv = __ldg(lst + (size_t)i * ESTRIDE + 1);
                        nl = 0ull - (((unsigned long long)qv.y << 32) | qv.x);
                        nh = 0ull - (((unsigned long long)qv.w << 32) | qv.z);
                    }
                    for (unsigned int k = 0; k < e.w; ++k) {
                        const unsigned int cs = WIDE ? pool32[e.z + k] : (unsigned int)p.pool[e.z + k];
                        if (cs == kPad || (int)cs == best_idx) continue;
                        retire(cs, nl, nh);
                    }
                }
            }
        }
        __syncthreads();
        {
            const unsigned int nq = min(s_qn[qp], (unsigned int)kQueue);
            for (unsigned int q = warp; q < nq; q += 32) {                    // a warp per pooled row
                const uint4 row = s_q[q];
                unsigned long long nl = 0, nh = 0;
                if (AF) {
                    const uint4 qv = __ldg(lst + (size_t)row.z * ESTRIDE + 1);
                    nl = 0ull - (((unsigned long long)qv.y << 32) | qv.x);
                    nh = 0ull - (((unsigned long long)qv.w << 32) | qv.z);
                }
                for (unsigned int k = lane; k < row.y; k += 32) {
                    const unsigned int cs = WIDE ? pool32[row.x + k] : (unsigned int)p.pool[row.x + k];
                    if (cs == kPad || (int)cs == best_idx) continue;          // the pick's own gain was reset above
                    retire(cs, nl, nh);
                }
            }
        }
        cluster.sync();                               // all decrements of this step have landed before anybody scans again
    }
    cluster.sync();                                   // nobody writes into a peer's slice any more
    for (int li = tid; li < n_own; li += blockDim.x) {
        const int s = crank + CL * li;
        p.gain_cnt[s] = s_cnt[li];
        if (AF) { p.gain_lo[s] = s_lo[li]; p.gain_hi[s] = s_hi[li]; }
        p.mask[s] = s_mask[li];
    }
    if (crank == 0 && tid == 0) {
        st->step = step;
        st->tot = tot;
        st->stop = stop;
        st->winner = -1;
        st->regain = 0;
        st->recompact = recompact;
        if (want_light) st->tail_single |= 2u;
        if (REFT) st->tie_step = tie_fallback ? 1u : 0u;
        if (p.dbg) p.dbg[11] += 1;
    }
}

// shared-memory layout of select_listcluster_kernel; 0 when a slice does not fit one SM
int listcluster_layout(const SelParams &p, int CL, ListClusterCfg *cfg, size_t *smem_bytes)
{
    if (p.V >= 0xffffffffll || p.S <= 0) return 0;
    const size_t n = ((size_t)p.S + CL - 1) / CL;
    const size_t budget = 190 * 1024;                       // + 41 KB static (queue of pooled rows, exchange slots)
    for (int stage = 1; stage >= 0; --stage) {
        size_t off = 0;
        auto take = [&](size_t bytes) { const size_t o = off; off = (off + bytes + 15) / 16 * 16; return (int)o; };
        take(n * 4);
        cfg->off_lo = take(p.af ? n * 8 : 0);
        cfg->off_hi = take(p.af ? n * 8 : 0);
        cfg->off_mask = take(n);
        cfg->off_loff = take(stage ? n * 4 : 0);
        cfg->off_llen = take(stage ? n * 4 : 0);
        if (off > budget) continue;
        cfg->n_own_max = (int)n;
        cfg->stage_lists = stage;
        cfg->min_recompact = 1u << 16;
        cfg->off_rows = cfg->off_vals = 0;
        cfg->row_cap = 0;
        if (p.ref_ties && p.af) {
            // the replay's sort buffer takes what is left: 12 bytes per row, a power of two of rows
            unsigned int cap = 1u << 13;
            while (cap >= 1024u && off + (size_t)cap * 12 + 32 > budget) cap >>= 1;
            if (p.tie_row_cap) {                            // UTMOS_OPT_TIE_ROW_CAP: a smaller buffer (tests of the fall-back step)
                unsigned int want = 2;
                while (want * 2 <= p.tie_row_cap) want <<= 1;
                if (cap >= 1024u && want < cap) cap = want;
            }
            if ((cap >= 1024u || p.tie_row_cap) && cap >= 2u && n <= 8192) {   // (the candidate slots alias the 32 KB queue of pooled rows)
                cfg->off_rows = take((size_t)cap * 4);
                cfg->off_vals = take((size_t)cap * 8);
                cfg->row_cap = cap;
            } else if (stage) {
                continue;                                   // try again without the staged list positions
            }
        }
        *smem_bytes = off;
        return 1;
    }
    return 0;
}

// CL: CTAs that share the per-sample state (1 = everything in one CTA)
int tail_layout(const SelParams &p, int CL, TailCfg *cfg, size_t *smem_bytes)
{
    if (p.V >= 0xffffffffll || p.S <= 0) return 0;          // rows are uint32 in the edge lists
    if (p.S > 65535 && CL == 1) return 0;                   // 16-bit carriers; wide cohorts need the cluster flavour
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = (off + bytes + 15) / 16 * 16; return (int)o; };
    const size_t S = ((size_t)p.S + CL - 1) / CL;           // samples whose state one CTA keeps
    take(S * 4);
    cfg->off_lo = take(p.af ? S * 8 : 0);
    cfg->off_hi = take(p.af ? S * 8 : 0);
    cfg->off_w = take(p.weights ? S * 8 : 0);
    cfg->off_mask = take(S);
    cfg->off_loff = take(CL == 1 ? S * 4 : 0);              // cluster flavour: list positions travel with the winners
    cfg->off_llen = take(CL == 1 ? S * 4 : 0);
    const size_t budget = 225 * 1024;
    if (off + 8 * 1024 > budget) return 0;
    const size_t live_bytes = (size_t)p.colPitchW * 4;
    cfg->live_words = 0;
    cfg->off_live = (int)off;
    // the live mask goes to shared memory when it leaves at least 16 KB for the staging area
    if (off + live_bytes + 16 * 1024 <= budget) { cfg->off_live = take(live_bytes); cfg->live_words = (int)p.colPitchW; }
    const size_t chunk = p.af ? 32 : 16;
    size_t stage = std::min<size_t>(budget - off, 64 * 1024) / chunk;
    cfg->off_stage = take(stage * chunk);
    cfg->stage_cap = (unsigned int)stage;
    cfg->min_recompact = 1u << 16;
    cfg->single_rows = 0;
    cfg->live_priv = nullptr;
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
    list_offsets_kernel<<<1, 1024, 0, stream>>>(p.gain_cnt, p.mask, p.S, list_off, list_len, cursor);
    *n_launch += 1;
    EdgeDst d;
    memset(&d, 0, sizeof(d));
    d.world = 1;
    d.lists[0] = lists;
    d.pool[0] = pool;
    d.slot_base = list_off;
    return launch_build_edges(stream, p, d, cursor, pool_cursor, n_launch);
}

// cursor[S] must be zero; entries go to d.lists[q][slot_base[s] + k] for every q < d.world
int launch_build_edges(cudaStream_t stream, const SelParams &p, const EdgeDst &d, unsigned int *cursor,
                       unsigned int *pool_cursor, int *n_launch)
{
    UT_CUDA(cudaMemsetAsync(pool_cursor, 0, 4, stream));
    long long blocks = (p.V + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    const bool wide = p.S > 65535;
    if (p.af && wide) build_edges_kernel<2, true><<<(unsigned)blocks, 256, 0, stream>>>(p, d, cursor, pool_cursor);
    else if (p.af) build_edges_kernel<2, false><<<(unsigned)blocks, 256, 0, stream>>>(p, d, cursor, pool_cursor);
    else if (wide) build_edges_kernel<1, true><<<(unsigned)blocks, 256, 0, stream>>>(p, d, cursor, pool_cursor);
    else build_edges_kernel<1, false><<<(unsigned)blocks, 256, 0, stream>>>(p, d, cursor, pool_cursor);
    *n_launch += 1;
    UT_CUDA(cudaGetLastError());
    return UTMOS_OK;
}

int launch_filter_lists(cudaStream_t stream, const SelParams &p, const uint4 *old_lists, const unsigned int *old_off,
                        const unsigned int *old_len, uint4 *new_lists, unsigned int *new_off, unsigned int *new_len,
                        int *n_launch)
{
    list_offsets_kernel<<<1, 1024, 0, stream>>>(p.gain_cnt, p.mask, p.S, new_off, new_len, nullptr);
    if (p.af) filter_edges_kernel<2><<<p.S, 256, 0, stream>>>(p.live, old_lists, old_off, old_len, new_lists, new_off);
    else filter_edges_kernel<1><<<p.S, 256, 0, stream>>>(p.live, old_lists, old_off, old_len, new_lists, new_off);
    *n_launch += 2;
    UT_CUDA(cudaGetLastError());
    return UTMOS_OK;
}

// CTAs of the tail kernel for this problem: 1 when the per-sample state fits one SM (and the caller does not ask for
// the owner-computes cluster flavour); otherwise the state is sliced over a cluster of 8, or 16, CTAs.  Wide cohorts
// (S > 65,535) always need the cluster.  0 = no configuration fits.
int tail_cluster_size(const SelParams &p, bool cluster)
{
    TailCfg cfg;
    size_t smem = 0;
    if (!cluster && tail_layout(p, 1, &cfg, &smem)) return 1;
    if (tail_layout(p, 8, &cfg, &smem)) return 8;
    if (tail_layout(p, 16, &cfg, &smem)) return 16;
    return 0;
}

// can the tail kernel run for S samples whatever the weights / live-mask size?  (multi-GPU: decided before the
// selection starts, identically on every rank)
int tail_possible(long long S, int af)
{
    SelParams p;
    memset(&p, 0, sizeof(p));
    p.S = (int)S;
    p.V = 1;
    p.af = af;
    p.colPitchW = 1ll << 30;                          // worst case: live mask in global memory
    static const double one = 1.0;
    p.weights = &one;                                 // worst case: weights present
    return tail_cluster_size(p, false) > 0;
}

int tail_plan(const SelParams &p, int *ok_out)
{
    // (the sample-major copy is not needed: the lists are built from the variant-major rows)
    *ok_out = p.rows != nullptr && p.S > 0 && tail_cluster_size(p, false) > 0;
    return UTMOS_OK;
}

// does the tail kernel keep the live mask of `p` in shared memory?  (else the cluster flavour needs one private copy
// of p.colPitchW words per CTA in global memory)
int tail_live_in_smem(const SelParams &p, bool cluster)
{
    TailCfg cfg;
    size_t smem = 0;
    const int CL = tail_cluster_size(p, cluster);
    return CL > 0 && tail_layout(p, CL, &cfg, &smem) && cfg.live_words > 0;
}

template <int ESTRIDE, bool FAST, int CL, bool WIDE>
static int launch_tail_t(cudaStream_t stream, const SelParams &p, const TailCfg &cfg, size_t smem, unsigned long long lists_total)
{
    auto kernel = select_tail_kernel<ESTRIDE, FAST, CL, WIDE>;
    UT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (CL == 1) {
        kernel<<<1, 1024, smem, stream>>>(p, cfg, lists_total);
        return UTMOS_OK;
    }
    if (CL > 8) UT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
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
    UT_CUDA(cudaLaunchKernelEx(&lc, kernel, p, cfg, lists_total));
    return UTMOS_OK;
}

template <int ESTRIDE, bool FAST, int CL, bool WIDE, bool REFT>
static int launch_listcluster_t(cudaStream_t stream, const SelParams &p, const ListClusterCfg &cfg, size_t smem,
                                unsigned long long lists_total, unsigned int light_rows)
{
    auto kernel = select_listcluster_kernel<ESTRIDE, FAST, CL, WIDE, REFT>;
    UT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (CL > 8) UT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
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
    UT_CUDA(cudaLaunchKernelEx(&lc, kernel, p, cfg, lists_total, light_rows));
    return UTMOS_OK;
}

// does the entry-divided cluster flavour exist for this problem (a slice of the per-sample state must fit one SM)?
bool listcluster_fits(const SelParams &p)
{
    ListClusterCfg cfg;
    size_t smem = 0;
    if (!listcluster_layout(p, 16, &cfg, &smem)) return false;
    return !(p.ref_ties && p.af) || (cfg.row_cap > 0 && p.af_vals != nullptr);
}

// The entry-divided cluster flavour of the tail (select_listcluster_kernel); it returns with bit 1 of st->tail_single set once a pick
// covers fewer than light_rows rows (0 = never) so that the single-SM tail can take over.
int launch_listcluster(cudaStream_t stream, const SelParams &p, unsigned long long lists_total, unsigned int light_rows,
                       int *n_launch)
{
    ListClusterCfg cfg;
    size_t smem = 0;
    if (!listcluster_layout(p, 16, &cfg, &smem)) { set_error("entry-divided tail: a slice of the state does not fit one SM"); return UTMOS_E_ARG; }
    const bool wide = p.S > 65535;
#define UT_LC(E, F, W, R) UT_TRY((launch_listcluster_t<E, F, 16, W, R>(stream, p, cfg, smem, lists_total, light_rows)))
    if (p.af && p.ref_ties) {
        if (!cfg.row_cap || !p.af_vals) { set_error("entry-divided tail: no room for the reference-tie replay"); return UTMOS_E_ARG; }
        if (wide) UT_LC(2, false, true, true); else UT_LC(2, false, false, true);
    }
    else if (p.af) { if (wide) UT_LC(2, false, true, false); else UT_LC(2, false, false, false); }
    else if (p.weights) { if (wide) UT_LC(1, false, true, false); else UT_LC(1, false, false, false); }
    else { if (wide) UT_LC(1, true, true, false); else UT_LC(1, true, false, false); }
#undef UT_LC
    *n_launch += 1;
    UT_CUDA(cudaGetLastError());
    return UTMOS_OK;
}

// cluster: the owner-computes cluster flavour (always for wide cohorts); it returns with st->tail_single set once a
// pick covers fewer than single_rows rows (0 = never hand over, the only choice for wide cohorts).
int launch_tail(cudaStream_t stream, const SelParams &p, unsigned long long lists_total, bool cluster,
                unsigned int single_rows, uint32_t *live_priv, int *n_launch)
{
    TailCfg cfg;
    size_t smem = 0;
    const bool wide = p.S > 65535;
    const int CL = tail_cluster_size(p, cluster);
    if (CL == 0 || !tail_layout(p, CL, &cfg, &smem)) { set_error("tail kernel: state does not fit in shared memory"); return UTMOS_E_ARG; }
    // hand-over to the single-CTA flavour only when that flavour exists for this problem
    cfg.single_rows = (CL > 1 && tail_cluster_size(p, false) == 1) ? single_rows : 0u;
    cfg.live_priv = live_priv;
    if (CL > 1 && cfg.live_words == 0 && !live_priv) { set_error("tail kernel: cluster flavour needs private live masks"); return UTMOS_E_ARG; }
#define UT_TAIL(E, F, C, W) UT_TRY((launch_tail_t<E, F, C, W>(stream, p, cfg, smem, lists_total)))
    if (wide) {
        if (CL == 16) { if (p.af) UT_TAIL(2, false, 16, true); else if (p.weights) UT_TAIL(1, false, 16, true); else UT_TAIL(1, true, 16, true); }
        else { if (p.af) UT_TAIL(2, false, 8, true); else if (p.weights) UT_TAIL(1, false, 8, true); else UT_TAIL(1, true, 8, true); }
    } else if (CL == 16) {
        if (p.af) UT_TAIL(2, false, 16, false); else if (p.weights) UT_TAIL(1, false, 16, false); else UT_TAIL(1, true, 16, false);
    } else if (CL == 8) {
        if (p.af) UT_TAIL(2, false, 8, false); else if (p.weights) UT_TAIL(1, false, 8, false); else UT_TAIL(1, true, 8, false);
    } else {
        if (p.af) UT_TAIL(2, false, 1, false);
        else if (p.weights) UT_TAIL(1, false, 1, false);
        else UT_TAIL(1, true, 1, false);
    }
#undef UT_TAIL
    *n_launch += 1;
    UT_CUDA(cudaGetLastError());
    return UTMOS_OK;
}

}  // namespace utmos
