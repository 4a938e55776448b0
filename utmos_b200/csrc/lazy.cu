// lazy.cu -- the greedy loop as LAZY greedy (Minoux), run by ONE CTA out of shared memory.
//
// Gains only fall while samples are picked (coverage is submodular), so a gain computed earlier is an upper bound
// now.  Instead of keeping every gain exact after every pick -- one decrement per (newly covered row, carrier) pair,
// 81.7 M of them on the 1kGP shape, bound by the atomic throughput of whoever applies them -- this kernel keeps
// per-sample ROW LISTS (the live rows that carry the sample, 4 bytes each) and re-counts only the few samples that can
// still win:
//
//   bound[s]   = length of s's list (every entry was live when the list was last compacted): an upper bound of the
//                gain, and the exact gain right after an evaluation
//   evaluate s = walk s's list, keep the entries whose row is still live (in place), bound[s] = what is left
//   a round    = every warp proposes its best not-yet-evaluated sample; those close to the best bound are evaluated,
//                all at once (one warp, or a group of warps, per candidate: one dependent global read per round);
//   a pick     = the best evaluated sample E, as soon as E beats every bound that is still unevaluated
//                (larger score, or equal score and lower index: np.argmax order, utmos/select.py:48).  E's list then
//                holds exactly the rows it newly covers: their live bits are cleared and every other evaluation
//                of this step is void again.
//
// The pick order is exactly the greedy order of utmos/select.py:24-53 / :91-112: a sample is only picked when its exact
// score is >= every other sample's upper bound, with the first-index rule applied to bounds as well.  The cost of a
// step no longer depends on how many carriers the covered rows have, only on the length of the lists looked at.
// Scores: count mode = list length; --weights = length * w (one float64 multiply of an exact integer); --af = exact
// fixed-point sum of the rows' AF limbs, rounded once (fixed_to_double), * w.  Weights must be >= 0 (a negative
// weight turns an upper bound of the gain into a lower bound of the score; such selections use the list-driven tail
// of tail.cu instead).
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace utmos {

namespace {

constexpr int kThreadsL = 1024;
constexpr int kNoIdx = 0x7fffffff;
constexpr unsigned int kNoRow = 0xffffffffu;

struct LKey {
    unsigned int hi, lo;     // order-preserving key of the score (count mode: hi = list length, lo = 0)
    int idx;
};

__device__ __forceinline__ bool lkey_better(const LKey &a, const LKey &b)
{
    return a.hi > b.hi || (a.hi == b.hi && (a.lo > b.lo || (a.lo == b.lo && a.idx < b.idx)));
}

template <int MODE>
__device__ __forceinline__ LKey warp_best(LKey k)
{
    LKey out;
    out.hi = __reduce_max_sync(0xffffffffu, k.hi);
    if (MODE == 0) {
        out.lo = 0u;
        out.idx = (int)__reduce_min_sync(0xffffffffu, k.hi == out.hi ? (unsigned int)k.idx : (unsigned int)kNoIdx);
    } else {
        out.lo = __reduce_max_sync(0xffffffffu, k.hi == out.hi ? k.lo : 0u);
        out.idx = (int)__reduce_min_sync(0xffffffffu, (k.hi == out.hi && k.lo == out.lo) ? (unsigned int)k.idx : (unsigned int)kNoIdx);
    }
    return out;
}

__device__ __forceinline__ unsigned long long dkey(double s)
{
    s += 0.0;                                   // -0.0 -> +0.0
    const unsigned long long b = (unsigned long long)__double_as_longlong(s);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__device__ __forceinline__ double dkey_score(unsigned int hi, unsigned int lo)
{
    const unsigned long long k = ((unsigned long long)hi << 32) | lo;
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

__device__ __forceinline__ unsigned int ld_cg_u32(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void group_barrier(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// one CTA scan: rl_off = exclusive scan of (mask == 1 ? gain_cnt : 0); rl_len = the same counts
__global__ void __launch_bounds__(1024) rowlist_offsets_kernel(const unsigned int *gain_cnt, const uint8_t *mask, int S,
                                                               unsigned int *rl_off, unsigned int *rl_len,
                                                               unsigned long long *total_out)
{
    __shared__ unsigned int s_warp[32];
    __shared__ unsigned long long s_carry;
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
        const unsigned long long carry = s_carry;
        if (i < S) {
            rl_off[i] = (unsigned int)(carry + s_warp[warp] + incl - v);
            rl_len[i] = v;
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = carry + s_warp[warp] + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = s_carry;
}

// Row list of sample s = the rows r with cols[s][r] & live[r], in any order (the list is a set).  One CTA per
// selectable sample streams its sample-major row once; a warp takes 32 words at a time, reserves room for their set
// bits with one shared-memory atomic and writes the row ids.
__global__ void __launch_bounds__(256) build_rowlists_kernel(const uint32_t *__restrict__ cols, const uint32_t *__restrict__ live,
                                                             const uint8_t *__restrict__ mask, long long colPitchW, long long V,
                                                             const unsigned int *__restrict__ rl_off, unsigned int *__restrict__ rl)
{
    __shared__ unsigned int s_cursor;
    const int s = blockIdx.x;
    if (mask[s] != 1) return;
    if (threadIdx.x == 0) s_cursor = 0;
    __syncthreads();
    const uint32_t *col = cols + (size_t)s * (size_t)colPitchW;
    unsigned int *dst = rl + rl_off[s];
    const long long words = (V + 31) >> 5;
    const int lane = threadIdx.x & 31;
    const long long wstep = blockDim.x;
    for (long long w0 = (threadIdx.x & ~31); w0 < words; w0 += wstep) {
        const long long w = w0 + lane;
        uint32_t x = 0u;
        if (w < words) x = __ldg(col + w) & __ldg(live + w);
        const unsigned int any = __ballot_sync(0xffffffffu, x != 0u);
        if (!any) continue;
        const int c = __popc(x);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        unsigned int base = 0;
        if (lane == 31) base = atomicAdd(&s_cursor, (unsigned int)total);
        base = __shfl_sync(0xffffffffu, base, 31);
        unsigned int pos = base + (unsigned int)(incl - c);
        while (x) {
            dst[pos++] = (unsigned int)(w << 5) + (unsigned int)(__ffs(x) - 1);
            x &= x - 1;
        }
    }
}

struct LazyCfg {
    int off_len, off_off, off_flag, off_w, off_lo, off_hi, off_live;   // byte offsets into dynamic shared memory
    int live_words;          // > 0: live mask in shared memory; 0: the global copy (p.live) is used in place
    int fresh;               // 1: the lists were just built -> every bound is exact
    int slack_shift;         // candidates with bound >= best - (best >> slack_shift) - 1 are evaluated together
    unsigned int g1_max;     // longest list one warp evaluates alone; longer ones take a group of 4 warps / the whole CTA
    unsigned int g4_max;
};

// flags: bit 0 = selectable (mask == 1), bit 1 = bound is exact (evaluated since the last pick)
template <int MODE>
__global__ void __launch_bounds__(kThreadsL, 1) select_lazy_kernel(SelParams p, LazyCfg cfg, unsigned int *rl,
                                                                   const unsigned int *rl_off, unsigned int *rl_len)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint4 s_we[2][32], s_wn[2][32];          // per-warp best evaluated / best unevaluated sample, by round parity
    __shared__ unsigned int s_tot[2][32];               // group evaluation: live entries of each warp's share of a batch
    __shared__ unsigned long long s_acc[8][2];          // group evaluation, AF: limb sums per group
    constexpr bool AF = MODE == 2;
    unsigned int *s_len = reinterpret_cast<unsigned int *>(smem + cfg.off_len);
    unsigned int *s_off = reinterpret_cast<unsigned int *>(smem + cfg.off_off);
    uint8_t *s_flag = smem + cfg.off_flag;
    double *s_w = reinterpret_cast<double *>(smem + cfg.off_w);
    unsigned long long *s_lo = reinterpret_cast<unsigned long long *>(smem + cfg.off_lo);
    unsigned long long *s_hi = reinterpret_cast<unsigned long long *>(smem + cfg.off_hi);
    uint32_t *lv = cfg.live_words > 0 ? reinterpret_cast<uint32_t *>(smem + cfg.off_live) : p.live;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = p.S;
    const bool has_w = MODE != 0 && p.weights != nullptr;
    SelState *st = p.st;

    for (int i = tid; i < S; i += kThreadsL) {
        const bool sel = p.mask[i] == 1;
        s_len[i] = sel ? rl_len[i] : 0u;
        s_off[i] = rl_off[i];
        s_flag[i] = (uint8_t)((sel ? 1 : 0) | ((sel && cfg.fresh) ? 2 : 0));
        if (has_w) s_w[i] = p.weights[i];
        if (AF) { s_lo[i] = p.gain_lo[i]; s_hi[i] = p.gain_hi[i]; }
    }
    for (int i = tid; i < cfg.live_words; i += kThreadsL) lv[i] = p.live[i];
    long long step = st->step, tot = st->tot;
    const long long limit = st->limit;
    int stop = st->stop;
    bool after_pick = !cfg.fresh;           // the next scan voids every evaluation
    int par = 0;
    unsigned long long n_walk = 0, n_cover = 0, n_round = 0, n_eval = 0;
    __syncthreads();

    auto make_key = [&](int i) {
        LKey k;
        k.idx = i;
        if (MODE == 0) {
            k.hi = s_len[i];
            k.lo = 0u;
        } else {
            double g = AF ? fixed_to_double(s_lo[i], s_hi[i], p.L, p.scale) : (double)s_len[i];
            if (has_w) g *= s_w[i];
            const unsigned long long kk = dkey(g);
            k.hi = (unsigned int)(kk >> 32);
            k.lo = (unsigned int)kk;
        }
        return k;
    };
    auto key_is_zero = [&](const LKey &k) { return MODE == 0 ? k.hi == 0u : dkey_score(k.hi, k.lo) == 0.0; };

    while (stop == 0 && step < limit) {
        // ---- scan: best evaluated (E) and best unevaluated (B) selectable sample, np.argmax order
        LKey e{0u, 0u, kNoIdx}, n{0u, 0u, kNoIdx};
        for (int i = tid; i < S; i += kThreadsL) {
            uint8_t f = s_flag[i];
            if (!(f & 1)) continue;
            if (after_pick && (f & 2)) { f &= (uint8_t)~2; s_flag[i] = f; }
            const LKey k = make_key(i);
            if (f & 2) { if (lkey_better(k, e)) e = k; }
            else { if (lkey_better(k, n)) n = k; }
        }
        after_pick = false;
        e = warp_best<MODE>(e);
        n = warp_best<MODE>(n);
        if (lane == 0) {
            s_we[par][warp] = make_uint4(e.hi, e.lo, (unsigned int)e.idx, 0u);
            s_wn[par][warp] = make_uint4(n.hi, n.lo, (unsigned int)n.idx, 0u);
        }
        __syncthreads();
        const uint4 qe = s_we[par][lane], qn = s_wn[par][lane];
        par ^= 1;
        const LKey we{qe.x, qe.y, (int)qe.z}, wn{qn.x, qn.y, (int)qn.z};      // lane l: the proposals of warp l
        const LKey E = warp_best<MODE>(we), B = warp_best<MODE>(wn);
        const bool haveE = E.idx != kNoIdx, haveB = B.idx != kNoIdx;
        if (haveE && (!haveB || lkey_better(E, B))) {
            // ---- E beats every upper bound: it is the greedy pick (utmos/select.py:48)
            if (key_is_zero(E)) { stop = UTMOS_STOP_ZERO; break; }       // utmos/select.py:51-52
            const int c = E.idx;
            const unsigned int cnt = s_len[c], off = s_off[c];
            if (tid == 0) {
                p.out_idx[step] = c;
                p.out_new[step] = cnt;
                p.out_score[step] = MODE == 0 ? (double)cnt : dkey_score(E.hi, E.lo);
                if (p.dbg_time) p.out_time[step] = global_timer_ns();
                p.mask[c] = 0;                                           // utmos/select.py:100
            }
            step += 1;
            tot += cnt;
            if (tot >= p.V) { stop = UTMOS_STOP_ALL; break; }            // utmos/select.py:110-112
            __syncthreads();                  // every thread has read s_len[c] / s_off[c] before the owner resets them
            if ((c & (kThreadsL - 1)) == tid) {
                s_flag[c] = 0;
                s_len[c] = 0;
                if (AF) { s_lo[c] = 0; s_hi[c] = 0; }
            }
            // its list holds exactly the rows it newly covers (it was compacted when E was evaluated): clear their bits
            for (unsigned int i = tid; i < cnt; i += kThreadsL) {
                const unsigned int r = ld_cg_u32(rl + off + i);
                atomicAnd(lv + (r >> 5), ~(1u << (r & 31)));
            }
            if (tid == 0) n_cover += cnt;
            after_pick = true;
            continue;                         // the barrier after the next scan orders the cleared bits before any evaluation
        }
        if (!haveB || (key_is_zero(B) && (!haveE || key_is_zero(E)))) { stop = UTMOS_STOP_ZERO; break; }
        // ---- evaluation round.  Lane l decides for warp l's proposal: worth evaluating = can still beat E and is
        // close to the best bound (anything else waits for a later round; correctness does not depend on the choice)
        bool part = wn.idx != kNoIdx && (!haveE || lkey_better(wn, E));
        if (part) {
            if (MODE == 0) {
                const unsigned int slack = (B.hi >> cfg.slack_shift) + 1u;
                part = wn.hi + slack >= B.hi;
            } else {
                const double sb = dkey_score(B.hi, B.lo), sn = dkey_score(wn.hi, wn.lo);
                part = sn >= sb - ldexp(sb, -cfg.slack_shift);
            }
        }
        const unsigned int my_len = part ? s_len[wn.idx] : 0u;
        const unsigned int lmax = __reduce_max_sync(0xffffffffu, my_len);
        const int G = lmax <= cfg.g1_max ? 1 : (lmax <= cfg.g4_max ? 4 : 32);
        const unsigned int pmask = __ballot_sync(0xffffffffu, part);
        if (tid == 0) n_round += 1;
        // which candidate does my group take?  G == 1: warp w its own proposal.  G > 1: the group g takes the
        // participant with rank g (best first)
        int cand = kNoIdx;
        const int group = warp / G, wig = warp % G;
        if (G == 1) {
            if ((pmask >> warp) & 1u) cand = __shfl_sync(0xffffffffu, wn.idx, warp);
        } else {
            int rank = 0;
            for (int j = 0; j < 32; ++j) {
                if (!((pmask >> j) & 1u)) continue;                       // warp-uniform
                LKey o;
                o.hi = __shfl_sync(0xffffffffu, wn.hi, j);
                o.lo = __shfl_sync(0xffffffffu, wn.lo, j);
                o.idx = __shfl_sync(0xffffffffu, wn.idx, j);
                if (lkey_better(o, wn)) rank += 1;
            }
            const unsigned int who = __ballot_sync(0xffffffffu, part && rank == group);
            if (who) cand = __shfl_sync(0xffffffffu, wn.idx, __ffs(who) - 1);
        }
        if (G > 1 && G < 32 && wig == 0 && lane == 0) { s_acc[group][0] = 0ull; s_acc[group][1] = 0ull; }
        if (G == 32 && tid == 0) { s_acc[0][0] = 0ull; s_acc[0][1] = 0ull; }
        if (cand != kNoIdx) {
            if (wig == 0 && lane == 0) n_eval += 1;
            const unsigned int off = s_off[cand], len = s_len[cand];
            unsigned int *lst = rl + off;
            const int GT = 32 * G, gl = wig * 32 + lane;
            unsigned int wbase = 0;
            unsigned long long a_lo = 0ull, a_hi = 0ull;
            constexpr int U = 8;
            int bpar = 0;
            for (unsigned int base = 0; base < len; base += (unsigned int)(GT * U)) {
                unsigned int r[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const unsigned int i = base + (unsigned int)(u * GT + gl);
                    r[u] = i < len ? ld_cg_u32(lst + i) : kNoRow;
                }
                unsigned int m[U];
                unsigned int mine = 0;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    bool alive = false;
                    if (r[u] != kNoRow) {
                        alive = (*reinterpret_cast<volatile uint32_t *>(lv + (r[u] >> 5)) >> (r[u] & 31)) & 1u;
                        n_walk += 1;
                    }
                    if (AF && alive) { a_lo += __ldg(p.q_lo + r[u]); a_hi += __ldg(p.q_hi + r[u]); }
                    m[u] = __ballot_sync(0xffffffffu, alive);
                    mine += (unsigned int)__popc(m[u]);
                }
                unsigned int my0 = wbase, batch = mine;
                if (G > 1) {
                    // all entries of the batch are in registers everywhere once the totals are exchanged: the survivors may
                    // then overwrite the front of the list (positions below base + GT*U only)
                    if (lane == 0) s_tot[bpar][warp] = mine;
                    if (G == 32) __syncthreads(); else group_barrier(1 + group, GT);
                    const unsigned int t = lane < G ? s_tot[bpar][group * G + lane] : 0u;
                    unsigned int incl = t;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const unsigned int x = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += x;
                    }
                    batch = __shfl_sync(0xffffffffu, incl, 31);
                    my0 = wbase + __shfl_sync(0xffffffffu, incl - t, wig);
                    bpar ^= 1;
                }
                unsigned int run = my0;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if ((m[u] >> lane) & 1u) lst[run + (unsigned int)__popc(m[u] & ((1u << lane) - 1u))] = r[u];
                    run += (unsigned int)__popc(m[u]);
                }
                wbase += batch;
            }
            if (AF) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    a_lo += __shfl_xor_sync(0xffffffffu, a_lo, o);
                    a_hi += __shfl_xor_sync(0xffffffffu, a_hi, o);
                }
                if (G > 1 && lane == 0) {
                    atomicAdd(&s_acc[G == 32 ? 0 : group][0], a_lo);
                    atomicAdd(&s_acc[G == 32 ? 0 : group][1], a_hi);
                }
            }
            if (G > 1 && AF) { if (G == 32) __syncthreads(); else group_barrier(1 + group, GT); }
            if (wig == 0 && lane == 0) {
                s_len[cand] = wbase;
                s_flag[cand] = 3;
                if (AF) {
                    s_lo[cand] = G > 1 ? s_acc[G == 32 ? 0 : group][0] : a_lo;
                    s_hi[cand] = G > 1 ? s_acc[G == 32 ? 0 : group][1] : a_hi;
                }
            }
        }
        __syncthreads();
    }

    __syncthreads();
    for (int i = tid; i < S; i += kThreadsL) {
        if (s_flag[i] & 1) {
            rl_len[i] = s_len[i];
            p.gain_cnt[i] = s_len[i];                  // upper bounds from here on (utmos_debug_gains recomputes)
            if (AF) { p.gain_lo[i] = s_lo[i]; p.gain_hi[i] = s_hi[i]; }
        } else {
            rl_len[i] = 0u;
        }
    }
    for (int i = tid; i < cfg.live_words; i += kThreadsL) p.live[i] = lv[i];
    if (tid == 0) {
        st->step = step;
        st->tot = tot;
        st->stop = stop;
        st->winner = -1;
        st->regain = 0;
        st->recompact = 0;
    }
    if (p.dbg) {                               // work counters: entries evaluated, rows covered, evaluation rounds, evaluations
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            n_walk += __shfl_xor_sync(0xffffffffu, n_walk, o);
            n_cover += __shfl_xor_sync(0xffffffffu, n_cover, o);
            n_round += __shfl_xor_sync(0xffffffffu, n_round, o);
            n_eval += __shfl_xor_sync(0xffffffffu, n_eval, o);
        }
        if (lane == 0) {
            atomicAdd(reinterpret_cast<unsigned long long *>(p.dbg + 5), n_walk);
            atomicAdd(reinterpret_cast<unsigned long long *>(p.dbg + 6), n_cover);
            atomicAdd(reinterpret_cast<unsigned long long *>(p.dbg + 7), n_round);
            atomicAdd(reinterpret_cast<unsigned long long *>(p.dbg + 12), n_eval);
            if (tid == 0) atomicAdd(reinterpret_cast<unsigned long long *>(p.dbg + 11), 1ull);
        }
    }
}

// shared-memory layout; returns 0 when the per-sample state does not fit one SM
int lazy_layout(const SelParams &p, bool weights, LazyCfg *cfg, size_t *smem_bytes)
{
    if (p.S <= 0 || p.V >= 0xfffffff0ll) return 0;
    const size_t S = (size_t)p.S;
    const size_t budget = 227 * 1024 - 4 * 1024;          // static shared memory of the kernel stays below 4 KB
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t at = off; off += (bytes + 15) / 16 * 16; return (int)at; };
    memset(cfg, 0, sizeof(*cfg));
    if (p.af) { cfg->off_lo = take(S * 8); cfg->off_hi = take(S * 8); }
    if (weights) cfg->off_w = take(S * 8);
    cfg->off_len = take(S * 4);
    cfg->off_off = take(S * 4);
    cfg->off_flag = take(S);
    if (off > budget) return 0;
    const size_t words = (size_t)((p.V + 31) >> 5);
    if (off + words * 4 <= budget) {
        cfg->off_live = take(words * 4);
        cfg->live_words = (int)words;
    }
    *smem_bytes = off;
    return 1;
}

}  // namespace

int lazy_possible(const SelParams &p, bool weights)
{
    LazyCfg cfg;
    size_t smem = 0;
    return p.cols != nullptr && lazy_layout(p, weights, &cfg, &smem);
}

// rl_off / rl_len from the (exact) gains of the selectable samples; *d_total = entries needed
int launch_rowlist_offsets(cudaStream_t stream, const SelParams &p, unsigned int *rl_off, unsigned int *rl_len,
                           unsigned long long *d_total, int *n_launch)
{
    rowlist_offsets_kernel<<<1, 1024, 0, stream>>>(p.gain_cnt, p.mask, p.S, rl_off, rl_len, d_total);
    *n_launch += 1;
    UT_CUDA(cudaGetLastError());
    return UTMOS_OK;
}

int launch_build_rowlists(cudaStream_t stream, const SelParams &p, const unsigned int *rl_off, unsigned int *rl,
                          int *n_launch)
{
    build_rowlists_kernel<<<p.S, 256, 0, stream>>>(p.cols, p.live, p.mask, p.colPitchW, p.V, rl_off, rl);
    *n_launch += 1;
    UT_CUDA(cudaGetLastError());
    return UTMOS_OK;
}

int launch_lazy(cudaStream_t stream, const SelParams &p, unsigned int *rl, const unsigned int *rl_off,
                unsigned int *rl_len, bool fresh, const int *tune, int *n_launch)
{
    LazyCfg cfg;
    size_t smem = 0;
    const bool weights = p.weights != nullptr;
    if (!lazy_layout(p, weights, &cfg, &smem)) { set_error("lazy kernel: state does not fit in shared memory"); return UTMOS_E_ARG; }
    cfg.fresh = fresh ? 1 : 0;
    // tune[0]: longest list one warp evaluates alone, tune[1]: ... a group of four warps (longer: the whole CTA),
    // tune[2]: candidates whose bound is within best >> tune[2] of the best bound are evaluated in the same round
    cfg.g1_max = (unsigned int)std::max(32, tune[0]);
    cfg.g4_max = std::max(cfg.g1_max, (unsigned int)std::max(32, tune[1]));
    cfg.slack_shift = std::max(0, std::min(31, tune[2]));
#define UT_LAZY(M)                                                                                                    \
    do {                                                                                                              \
        UT_CUDA(cudaFuncSetAttribute(select_lazy_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        select_lazy_kernel<M><<<1, kThreadsL, smem, stream>>>(p, cfg, rl, rl_off, rl_len);                            \
    } while (0)
    if (p.af) UT_LAZY(2);
    else if (weights) UT_LAZY(1);
    else UT_LAZY(0);
#undef UT_LAZY
    *n_launch += 1;
    UT_CUDA(cudaGetLastError());
    return UTMOS_OK;
}

}  // namespace utmos
