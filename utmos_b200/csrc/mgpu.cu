// mgpu.cu -- multi-GPU hand-over from the row-sharded head to the replicated tail.
//
// The greedy loop is latency bound once the common variants are gone (tail.cu): a single SM running from shared
// memory is faster than any scheme that synchronises GPUs every step.  So the ranks share the work only while
// it is bandwidth bound (ingest, transpose, column reduce, the head picks with their per-step delta exchange,
// select.cu: select_mgpu_kernel).  When the live part of the matrix has become sparse, every rank turns ITS
// live rows into edge-list entries and stores them straight into the merged per-sample lists of EVERY rank over
// NVLink (peer pointers of the IPC-mapped exchange blocks); all ranks then hold byte-identical lists, live
// mask and gains and run the same deterministic single-CTA tail kernel -- no further communication, and every
// rank returns the same report rows.
//
//   1. live_colpop_kernel     lcnt[s] = live rows of this rank that carry sample s
//   2. gather_counts_kernel   all-gather of lcnt through the peers' inboxes (+ sequence flags)
//   3. gather_offsets_kernel  merged list directory: list_off / list_len (replicated) and this rank's first
//                             slot per sample (my_base), its pool share
//   4. build_edges_kernel     (tail.cu) into this rank's own merged buffers, then gather_push_*_kernel copy its
//                             contiguous segments / pool share to every peer  +  gather_live_kernel (live mask)
//   5. gather_done_kernel     system-scope fence, completion flags, wait for every peer
#include "common.cuh"

namespace utmos {

namespace {

constexpr long long kGatherSpinLimit = 1ll << 26;

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) live_colpop_kernel(const uint32_t *__restrict__ cols, const uint32_t *__restrict__ live,
                                                          const uint8_t *__restrict__ mask, long long colPitchW, int S,
                                                          unsigned int *out)
{
    const int s = blockIdx.x;
    if (s >= S) return;
    if (mask[s] != 1) {                        // no list for a sample that can never be picked (tail.cu: build_edges_kernel)
        if (threadIdx.x == 0) out[s] = 0u;
        return;
    }
    const uint4 *col = reinterpret_cast<const uint4 *>(cols + (long long)s * colPitchW);
    const uint4 *lv = reinterpret_cast<const uint4 *>(live);
    const long long n4 = colPitchW / 4;
    unsigned int alive = 0;
    for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
        const uint4 c = ld_stream_u128(col + i);
        const uint4 l = __ldcg(lv + i);
        alive += __popc(c.x & l.x) + __popc(c.y & l.y) + __popc(c.z & l.z) + __popc(c.w & l.w);
    }
    __shared__ unsigned int s_alive[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) alive += __shfl_xor_sync(0xffffffffu, alive, o);
    if ((threadIdx.x & 31) == 0) s_alive[threadIdx.x >> 5] = alive;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int a = 0;
        for (int i = 0; i < 8; ++i) a += s_alive[i];
        out[s] = a;
    }
}

// wait until every peer has published sequence number `seq` in my flags; sets st->abort_flag on timeout
__device__ __forceinline__ void wait_peers(const GatherParams &g)
{
    for (int q = 0; q < g.world; ++q) {
        if (q == g.rank) continue;
        long long spins = 0;
        while (ld_acquire_sys(g.flags + q) < g.seq) {
            if (++spins > kGatherSpinLimit) { atomicExch(&g.st->abort_flag, 3u); return; }
        }
    }
}

__global__ void __launch_bounds__(1024) gather_counts_kernel(GatherParams g)
{
    const size_t slot = ((size_t)(g.seq & 1) * g.world + g.rank) * (size_t)g.S;
    for (int i = threadIdx.x; i < g.S; i += blockDim.x) {
        const unsigned int v = g.lcnt[i];
        g.inbox_cnt[slot + i] = v;
        for (int q = 0; q < g.world; ++q)
            if (q != g.rank) g.peer_inbox_cnt[q][slot + i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < g.world && (int)threadIdx.x != g.rank) st_release_sys(g.peer_flags[threadIdx.x] + g.rank, g.seq);
    if (threadIdx.x == 0) wait_peers(g);
}

// list_len[s] = sum over ranks of their live counts; list_off = exclusive scan; my_base[s] = list_off[s] + counts of
// the ranks before this one; pool_base[0] = pool shares of the ranks before this one; cursor[s] = 0.
__global__ void __launch_bounds__(1024) gather_offsets_kernel(GatherParams g, unsigned int *list_off, unsigned int *list_len,
                                                              unsigned int *my_base, unsigned int *cursor,
                                                              unsigned int *pool_base)
{
    __shared__ unsigned int s_warp[32];
    __shared__ unsigned int s_carry;
    __shared__ unsigned long long s_tot[kMaxRanks][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t base0 = (size_t)(g.seq & 1) * g.world * (size_t)g.S;
    if (threadIdx.x == 0) s_carry = 0;
    unsigned long long tot[kMaxRanks];
#pragma unroll
    for (int q = 0; q < kMaxRanks; ++q) tot[q] = 0;
    __syncthreads();
    for (int base = 0; base < g.S; base += blockDim.x) {
        const int i = base + threadIdx.x;
        unsigned int v = 0, before = 0;
        if (i < g.S) {
#pragma unroll
            for (int q = 0; q < kMaxRanks; ++q) {
                if (q < g.world) {
                    const unsigned int c = __ldcv(g.inbox_cnt + base0 + (size_t)q * g.S + i);
                    v += c;
                    if (q < g.rank) before += c;
                    tot[q] += c;
                }
            }
        }
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
        if (i < g.S) {
            const unsigned int off = carry + s_warp[warp] + incl - v;
            list_off[i] = off;
            list_len[i] = v;
            my_base[i] = off + before;
            cursor[i] = 0;
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = carry + s_warp[warp] + incl;
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < kMaxRanks; ++q) {
        unsigned long long t = tot[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) s_tot[q][warp] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long pb = 0;
        for (int q = 0; q < g.rank; ++q) {
            unsigned long long t = 0;
            for (int w = 0; w < 32; ++w) t += s_tot[q][w];
            pb += (t * 2 + t / 4 + 64 + 7) & ~7ull;          // == mgpu_pool_share(t)
        }
        pool_base[0] = (unsigned int)pb;
    }
}

__global__ void __launch_bounds__(256) gather_live_kernel(GatherParams g, const uint32_t *__restrict__ live)
{
    const long long n = g.live_words;
    for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < n; w += (long long)gridDim.x * blockDim.x) {
        const uint32_t v = live[w];
        for (int q = 0; q < g.world; ++q) g.live_dst[q][g.live_word0 + w] = v;
    }
}

// The entries a rank built for sample s are contiguous in the merged layout ([my_base[s], + lcnt[s])), and so is its
// share of the pool.  Scattered 16-byte stores over NVLink are transaction bound (measured: 1.2 s for 250 M entries
// between two B200s), so build_edges_kernel writes into THIS rank's buffers only and these kernels then copy whole
// segments to the peers with coalesced 128-bit stores.
__global__ void __launch_bounds__(256) gather_push_lists_kernel(GatherParams g)
{
    const uint4 *mine = g.lists_dst[g.rank];
    for (int s = blockIdx.x; s < g.S; s += gridDim.x) {
        const size_t first = (size_t)g.my_base[s] * g.estride;
        const size_t n = (size_t)g.lcnt[s] * g.estride;
        for (int q = 0; q < g.world; ++q) {
            if (q == g.rank) continue;
            uint4 *dst = g.lists_dst[q] + first;
            for (size_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = mine[first + i];
        }
    }
}

__global__ void __launch_bounds__(256) gather_push_pool_kernel(GatherParams g)
{
    const size_t first = (size_t)(*g.pool_base) * g.pool_elem / 16;          // shares start on 16-byte boundaries
    const size_t n = ((size_t)(*g.pool_cursor) * g.pool_elem + 15) / 16;
    const uint4 *mine = reinterpret_cast<const uint4 *>(g.pool_dst[g.rank]) + first;
    for (int q = 0; q < g.world; ++q) {
        if (q == g.rank) continue;
        uint4 *dst = reinterpret_cast<uint4 *>(g.pool_dst[q]) + first;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = mine[i];
    }
}

__global__ void gather_done_kernel(GatherParams g)
{
    __threadfence_system();
    if ((int)threadIdx.x < g.world && (int)threadIdx.x != g.rank) st_release_sys(g.peer_flags[threadIdx.x] + g.rank, g.seq);
    __syncwarp();
    if (threadIdx.x == 0) wait_peers(g);
}

}  // namespace

// pool entries reserved for a rank that contributes `live_bits` list entries (rows with >= 7 carriers keep their
// carrier list padded to a multiple of 8: at most 15/7 entries per live bit)
unsigned long long mgpu_pool_share(unsigned long long live_bits) { return (live_bits * 2 + live_bits / 4 + 64 + 7) & ~7ull; }

int launch_live_counts(cudaStream_t stream, const SelParams &p, unsigned int *lcnt, int *n_launch)
{
    if (!p.cols || p.V <= 0) {
        UT_CUDA(cudaMemsetAsync(lcnt, 0, (size_t)p.S * 4, stream));
        return UTMOS_OK;
    }
    live_colpop_kernel<<<p.S, 256, 0, stream>>>(p.cols, p.live, p.mask, p.colPitchW, p.S, lcnt);
    *n_launch += 1;
    UT_CUDA(cudaGetLastError());
    return UTMOS_OK;
}

int launch_gather_counts(cudaStream_t stream, const GatherParams &g, int *n_launch)
{
    gather_counts_kernel<<<1, 1024, 0, stream>>>(g);
    *n_launch += 1;
    UT_CUDA(cudaGetLastError());
    return UTMOS_OK;
}

int launch_gather_offsets(cudaStream_t stream, const GatherParams &g, unsigned int *list_off, unsigned int *list_len,
                          unsigned int *my_base, unsigned int *cursor, unsigned int *pool_base, int *n_launch)
{
    gather_offsets_kernel<<<1, 1024, 0, stream>>>(g, list_off, list_len, my_base, cursor, pool_base);
    *n_launch += 1;
    UT_CUDA(cudaGetLastError());
    return UTMOS_OK;
}

int launch_gather_live(cudaStream_t stream, const GatherParams &g, const uint32_t *live, int *n_launch)
{
    if (g.live_words <= 0) return UTMOS_OK;
    long long blocks = (g.live_words + 255) / 256;
    if (blocks > 148 * 4) blocks = 148 * 4;
    gather_live_kernel<<<(unsigned)blocks, 256, 0, stream>>>(g, live);
    *n_launch += 1;
    UT_CUDA(cudaGetLastError());
    return UTMOS_OK;
}

int launch_gather_push(cudaStream_t stream, const GatherParams &g, int *n_launch)
{
    if (g.world <= 1) return UTMOS_OK;
    gather_push_lists_kernel<<<148 * 8, 256, 0, stream>>>(g);
    gather_push_pool_kernel<<<148 * 4, 256, 0, stream>>>(g);
    *n_launch += 2;
    UT_CUDA(cudaGetLastError());
    return UTMOS_OK;
}

int launch_gather_done(cudaStream_t stream, const GatherParams &g, int *n_launch)
{
    gather_done_kernel<<<1, 32, 0, stream>>>(g);
    *n_launch += 1;
    UT_CUDA(cudaGetLastError());
    return UTMOS_OK;
}

}  // namespace utmos
