// ingest.cu -- K2: raw chunk (packed .jl rows, dense bool rows or dense float32 GT*AF rows, already in HBM)
//              -> order-preserving compaction of the informative rows into the variant-major bit matrix.
//
// Replaces utmos/select.py:275-280 (np.unpackbits(count=S) -> .any(axis=1) -> boolean row filter on GT and
// AF) without ever unpacking: rows stay bit packed, only the bit order changes from np.packbits' MSB-first
// bytes to LSB-first 32-bit words (sample s = word s>>5, bit s&31) and the pitch is padded to 16 bytes so
// every later kernel can use aligned 128-bit loads.
//
// Algorithmic bytes per chunk: read V*P (+8V AF), write V'*pitch (+8V').
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace utmos {

namespace {

constexpr int kRowsPerBlock = 256;   // one thread per row in the scan, 8 warps move the rows
constexpr int kThreads = 256;

// natural-order word w of packed MSB-first row (lane-independent path)
__device__ __forceinline__ uint32_t packed_word(const uint8_t *row, int w, int nbytes, int S)
{
    const int b0 = w * 4;
    uint32_t le = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (b0 + k < nbytes) le |= (uint32_t)__ldg(row + b0 + k) << (8 * k);
    uint32_t out = msb_bytes_to_word(le);
    const int valid = S - w * 32;
    if (valid < 32) out &= valid <= 0 ? 0u : ((1u << valid) - 1u);
    return out;
}

// one warp builds word w of a dense row by ballot (coalesced 32-sample loads); every lane gets the word
template <int KIND>
__device__ __forceinline__ uint32_t dense_word(const void *raw, long long r, int w, int S, int lane, float *vmax)
{
    const int s = w * 32 + lane;
    bool on = false;
    if (KIND == RAW_DENSE_U8) {
        if (s < S) on = __ldg((const uint8_t *)raw + r * (long long)S + s) != 0;
    } else {
        if (s < S) {
            const float v = __ldg((const float *)raw + r * (long long)S + s);
            on = v != 0.0f;
            if (on && vmax) *vmax = fmaxf(*vmax, v);
        }
    }
    return __ballot_sync(0xffffffffu, on);
}

// k1: one warp per row -> flags[r] = row has any of the first S bits set; per-block kept counts.
// Float flavour also recovers the row's AF (all nonzero entries of a GT*AF row are equal, select.py:222).
template <int KIND>
__global__ void __launch_bounds__(kThreads) flag_rows_kernel(const void *raw, long long n_rows, long long pitch_in,
                                                             int S, uint8_t *flags, unsigned int *block_counts,
                                                             double *af_tmp)
{
    __shared__ unsigned int s_count;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nW = (S + 31) / 32;
    const int nbytes = (S + 7) / 8;
    const long long row0 = (long long)blockIdx.x * kRowsPerBlock;
    unsigned int mine = 0;
    for (int i = warp; i < kRowsPerBlock; i += kThreads / 32) {
        const long long r = row0 + i;
        if (r >= n_rows) break;
        uint32_t any = 0;
        float vmax = 0.0f;
        if (KIND == RAW_PACKED_MSB) {
            const uint8_t *row = (const uint8_t *)raw + r * pitch_in;
            for (int w = lane; w < nW; w += 32) any |= packed_word(row, w, nbytes, S);
            any = __ballot_sync(0xffffffffu, any != 0);
        } else {
            for (int w = 0; w < nW; ++w) any |= dense_word<KIND>(raw, r, w, S, lane, &vmax);
            if (KIND == RAW_DENSE_F32) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
            }
        }
        // dense (hdf5) rows are all kept: data.shape[0] of the file is num_vars (utmos/select.py:153); only
        // .jl parts go through the uninformative-row filter (utmos/select.py:276-280)
        if (KIND != RAW_PACKED_MSB) any = 1;
        if (lane == 0) {
            flags[r] = any ? 1 : 0;
            mine += any ? 1u : 0u;
            if (KIND == RAW_DENSE_F32) af_tmp[r] = (double)vmax;
        }
    }
    if (lane == 0 && mine) atomicAdd(&s_count, mine);
    __syncthreads();
    if (threadIdx.x == 0) block_counts[blockIdx.x] = s_count;
}

// exclusive scan across one CTA (blockDim.x multiple of 32, <= 1024); returns prefix, *total = CTA sum
__device__ __forceinline__ unsigned int block_exclusive_scan(unsigned int v, unsigned int *total)
{
    __shared__ unsigned int s_warp[32];
    __shared__ unsigned int s_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    unsigned int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned int w = lane < nwarp ? s_warp[lane] : 0u;
        unsigned int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        s_warp[lane] = wi - w;
        if (lane == 31) s_total = wi;
    }
    __syncthreads();
    const unsigned int out = incl - v + s_warp[warp];
    *total = s_total;
    __syncthreads();
    return out;
}

// k2: exclusive scan of the per-block kept counts (single CTA, running carry); d_chunk_total = sum
__global__ void __launch_bounds__(1024) scan_blocks_kernel(const unsigned int *block_counts,
                                                           unsigned int *block_offsets, long long n_blocks,
                                                           long long *d_chunk_total)
{
    long long carry = 0;
    for (long long base = 0; base < n_blocks; base += blockDim.x) {
        const long long i = base + threadIdx.x;
        const unsigned int v = i < n_blocks ? block_counts[i] : 0u;
        unsigned int total;
        const unsigned int ex = block_exclusive_scan(v, &total);
        if (i < n_blocks) block_offsets[i] = (unsigned int)(carry + ex);
        carry += total;
    }
    if (threadIdx.x == 0) *d_chunk_total = carry;
}

// k3: move the kept rows (bit order converted, pitch padded with zero words) behind the rows already stored
template <int KIND>
__global__ void __launch_bounds__(kThreads) scatter_rows_kernel(const void *raw, long long n_rows, long long pitch_in,
                                                                const double *af_in, int S, int pitchW,
                                                                const uint8_t *flags,
                                                                const unsigned int *block_offsets,
                                                                const long long *d_nrows, uint32_t *rows_out,
                                                                double *af_out)
{
    __shared__ long long s_dest[kRowsPerBlock];
    const long long row0 = (long long)blockIdx.x * kRowsPerBlock;
    const long long r_mine = row0 + threadIdx.x;
    const unsigned int keep = (r_mine < n_rows && flags[r_mine]) ? 1u : 0u;
    unsigned int total;
    const unsigned int ex = block_exclusive_scan(keep, &total);
    const long long base = *d_nrows + block_offsets[blockIdx.x];
    s_dest[threadIdx.x] = keep ? base + ex : -1;
    if (keep && af_out && af_in) af_out[base + ex] = af_in[r_mine];
    __syncthreads();
    if (total == 0) return;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nW = (S + 31) / 32;
    const int nbytes = (S + 7) / 8;
    for (int i = warp; i < kRowsPerBlock; i += kThreads / 32) {
        const long long dest = s_dest[i];
        if (dest < 0) continue;
        const long long r = row0 + i;
        uint32_t *out = rows_out + dest * pitchW;
        if (KIND == RAW_PACKED_MSB) {
            const uint8_t *row = (const uint8_t *)raw + r * pitch_in;
            for (int w = lane; w < pitchW; w += 32) out[w] = w < nW ? packed_word(row, w, nbytes, S) : 0u;
        } else {
            for (int w0 = 0; w0 < pitchW; w0 += 32) {
                uint32_t keepw = 0;
                for (int j = 0; j < 32; ++j) {
                    const int w = w0 + j;
                    if (w >= nW) break;                     // warp-uniform
                    const uint32_t x = dense_word<KIND>(raw, r, w, S, lane, nullptr);
                    if (lane == j) keepw = x;
                }
                if (w0 + lane < pitchW) out[w0 + lane] = keepw;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// fast path for packed .jl rows (the common case): ONE pass.  A CTA stages a tile of consecutive raw rows in
// shared memory with aligned 128-bit streaming loads (rows are contiguous, so the tile is one byte range
// whatever the row pitch), flags the informative rows, gets its output offset by a decoupled look-back over
// the tiles before it (tile order = ticket order, so the chain always makes progress), and writes the kept
// rows -- which are consecutive in the output -- as aligned 128-bit stores.
// ------------------------------------------------------------------------------------------------
constexpr int kFastMaxRows = 256;
constexpr size_t kFastSmemRaw = 64 * 1024;        // largest row pitch the fast path stages
constexpr size_t kFastTileBytes = 54 * 1024;      // preferred tile: 4 CTAs per SM; measured 0.209 ms against 0.231 ms with 27 KB tiles
                                                  // (fewer tiles = fewer look-backs; the TMA keeps the loads in flight)
constexpr unsigned long long kTileAgg = 1ull << 62, kTilePrefix = 2ull << 62, kTileValue = (1ull << 62) - 1ull;


__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// ---- TMA (1-D bulk copy) + mbarrier, raw PTX.  One thread arms the barrier with the byte count and issues the copies;
// everybody waits on the barrier's phase.  The staged tile is ONE contiguous byte run of the input whatever the row pitch.
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned int count)
{
    const unsigned int a = (unsigned int)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned int bytes)
{
    const unsigned int a = (unsigned int)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}

__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gmem_src, unsigned int bytes, unsigned long long *bar)
{
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
    const unsigned int b = (unsigned int)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(gmem_src), "r"(bytes), "r"(b) : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned int parity)
{
    const unsigned int a = (unsigned int)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(a), "r"(parity) : "memory");
}

// little-endian 4 bytes at byte offset o of the staged tile (any alignment)
__device__ __forceinline__ uint32_t smem_le32(const uint32_t *s32, unsigned int o)
{
    const unsigned int a = o >> 2;
    return __funnelshift_r(s32[a], s32[a + 1], (o & 3u) * 8u);
}

// bytes [lo, hi) of the staged tile, hi - lo < 16: nonzero?
__device__ __forceinline__ uint32_t smem_bytes_any(const uint32_t *s32, unsigned int lo, unsigned int hi)
{
    uint32_t any = 0;
    for (unsigned int j = lo; j < hi; j += 4) {
        uint32_t x = smem_le32(s32, j);
        const unsigned int n = hi - j;
        if (n < 4) x &= (1u << (8u * n)) - 1u;
        any |= x;
    }
    return any;
}

// FLAVOUR 0: one warp ORs the words of a row (first version).  FLAVOUR 1: the load loop records which 16-byte
// pieces are nonzero in a bitmap, one thread per row then tests the bitmap bits of the pieces that lie wholly
// inside its row and the few bytes it shares with its neighbours; the store loop splits its unit index without
// an integer division; the look-back can read kLook tiles per lane and round trip (window = 32 * kLook tiles).
// FLAVOUR 3: the tile is staged by the TMA (cp.async.bulk, one elected thread, mbarrier complete_tx) instead of
// LDG + STS in every thread, the nonzero-piece bitmap is then read back from shared memory, and the look-back reads
// 256 tile states per round trip (every warp of the CTA takes a window of 32) instead of 32.
template <int FLAVOUR>
__global__ void __launch_bounds__(kThreads) ingest_packed_kernel(const uint8_t *__restrict__ raw, long long n_rows,
                                                                 long long pitch_in, const double *__restrict__ af_in,
                                                                 int S, int pitchW, int rows_per_tile, long long n_tiles,
                                                                 unsigned long long *state, const long long *d_nrows,
                                                                 long long *d_chunk_total, uint32_t *__restrict__ rows_out,
                                                                 double *__restrict__ af_out)
{
    extern __shared__ __align__(16) uint8_t f_smem[];
    __shared__ unsigned int s_tile;
    __shared__ unsigned long long s_excl;
    __shared__ unsigned short s_src[kFastMaxRows];
    __shared__ uint8_t s_flag[kFastMaxRows];
    __shared__ uint32_t s_nz[(kFastSmemRaw + 64) / 16 / 32 + 10];      // FLAVOUR 1: bit i = staged piece i is nonzero
    __shared__ __align__(8) unsigned long long s_mbar;                 // FLAVOUR 3
    __shared__ unsigned long long s_lb_sum[kThreads / 32];
    __shared__ unsigned int s_lb_has[kThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (FLAVOUR == 3 && tid == 0) mbar_init(&s_mbar, 1u);
    if (tid == 0) s_tile = (unsigned int)atomicAdd(state, 1ull);
    __syncthreads();
    const long long tile = s_tile;
    if (tile >= n_tiles) return;
    const long long row0 = tile * rows_per_tile;
    const int nr = (int)min((long long)rows_per_tile, n_rows - row0);
    const long long total_bytes = n_rows * pitch_in;
    const long long start = row0 * pitch_in, end = start + (long long)nr * pitch_in;
    const long long a0 = start & ~15ll;
    const unsigned int mis = (unsigned int)(start - a0);
    const int n16 = (int)((end - a0 + 15) >> 4);
    uint4 *s16 = reinterpret_cast<uint4 *>(f_smem);
    const uint32_t *s32 = reinterpret_cast<const uint32_t *>(f_smem);
    const int nW = (S + 31) / 32;
    const uint32_t last_mask = (S & 31) ? ((1u << (S & 31)) - 1u) : 0xffffffffu;
    if (FLAVOUR == 0) {
        for (int i = tid; i < n16 + 2; i += kThreads) {        // two extra zeroed pieces: reads past the last row stay in bounds
            const long long off = a0 + 16ll * i;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (i < n16) {
                if (off + 16 <= total_bytes) {
                    v = ld_stream_u128(reinterpret_cast<const uint4 *>(raw + off));
                } else {
                    uint32_t w[4] = {0u, 0u, 0u, 0u};
                    for (int b = 0; b < 16 && off + b < total_bytes; ++b) w[b >> 2] |= (uint32_t)__ldg(raw + off + b) << (8 * (b & 3));
                    v = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            s16[i] = v;
        }
        __syncthreads();
        for (int i = warp; i < nr; i += kThreads / 32) {
            const unsigned int o = mis + (unsigned int)i * (unsigned int)pitch_in;
            uint32_t any = 0;
            for (int k = lane; k < nW; k += 32) {
                uint32_t x = smem_le32(s32, o + 4u * k);          // bit order does not matter for "any"
                if (k == nW - 1) x = msb_bytes_to_word(x) & last_mask;
                any |= x;
            }
            any = __ballot_sync(0xffffffffu, any != 0);
            if (lane == 0) s_flag[i] = any ? 1 : 0;
        }
    } else {
        const uint4 *g16 = reinterpret_cast<const uint4 *>(raw + a0);
        const int n_in = (int)min((long long)n16, (total_bytes - a0) >> 4);     // pieces that lie wholly inside the input
        const int n_iter = (n16 + 2 + kThreads - 1) / kThreads * kThreads;      // whole warps: the ballot needs every lane
        if (FLAVOUR == 3) {
            // n_in whole pieces come from the TMA; the ragged last piece of the input (and the two zero pieces behind the
            // tile) are written by ordinary stores to OTHER addresses
            if (tid == 0 && n_in > 0) {
                mbar_expect_tx(&s_mbar, (unsigned int)n_in * 16u);
                for (int c0 = 0; c0 < n_in; c0 += 1024) {                       // copies of at most 16 KB
                    const int n = min(1024, n_in - c0);
                    bulk_load(s16 + c0, g16 + c0, (unsigned int)n * 16u, &s_mbar);
                }
            }
            for (int i = n_in + tid; i < n16 + 2; i += kThreads) {
                uint4 x = make_uint4(0u, 0u, 0u, 0u);
                if (i < n16) {
                    const long long off = a0 + 16ll * i;
                    unsigned long long lo = 0ull, hi = 0ull;
                    for (int b = 0; b < 16 && off + b < total_bytes; ++b) {
                        const unsigned long long y = __ldg(raw + off + b);
                        if (b < 8) lo |= y << (8 * b);
                        else hi |= y << (8 * (b - 8));
                    }
                    x = make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
                }
                s16[i] = x;
            }
            __syncthreads();                                                    // ragged pieces visible, barrier initialised
            if (n_in > 0) mbar_wait(&s_mbar, 0u);
            for (int i = tid; i < n_iter; i += kThreads) {
                uint4 x = make_uint4(0u, 0u, 0u, 0u);
                if (i < n16 + 2) x = s16[i];
                const uint32_t nz = __ballot_sync(0xffffffffu, (x.x | x.y | x.z | x.w) != 0u);
                if (lane == 0 && i < n16 + 2) s_nz[i >> 5] = nz;
            }
        } else if (FLAVOUR == 2) {
            // opt-in: kLoadBatch loads of a thread are issued before the first of them is used.  In FLAVOUR 1 the vote
            // on the loaded piece sits in the load loop, so every trip waits for its own load (SASS: one LDG.128 and
            // one VOTE per trip) -- seven dependent DRAM round trips per tile.
            constexpr int kLoadBatch = 4;
            for (int i0 = tid; i0 < n_iter; i0 += kLoadBatch * kThreads) {
                uint4 v[kLoadBatch];
#pragma unroll
                for (int k = 0; k < kLoadBatch; ++k) {
                    const int i = i0 + k * kThreads;
                    v[k] = make_uint4(0u, 0u, 0u, 0u);
                    if (i < n_in) v[k] = ld_stream_u128(g16 + i);
                }
#pragma unroll
                for (int k = 0; k < kLoadBatch; ++k) {
                    const int i = i0 + k * kThreads;
                    if (i >= n_iter) break;                                     // the same for every thread of the CTA
                    uint4 x = v[k];
                    if (i >= n_in && i < n16) {                                 // the last piece of the input, byte by byte
                        const long long off = a0 + 16ll * i;
                        unsigned long long lo = 0ull, hi = 0ull;
                        for (int b = 0; b < 16 && off + b < total_bytes; ++b) {
                            const unsigned long long y = __ldg(raw + off + b);
                            if (b < 8) lo |= y << (8 * b);
                            else hi |= y << (8 * (b - 8));
                        }
                        x = make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
                    }
                    if (i < n16 + 2) s16[i] = x;
                    const uint32_t nz = __ballot_sync(0xffffffffu, (x.x | x.y | x.z | x.w) != 0u);
                    if (lane == 0 && i < n16 + 2) s_nz[i >> 5] = nz;
                }
            }
        } else {
            for (int i = tid; i < n_iter; i += kThreads) {
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (i < n_in) {
                    v = ld_stream_u128(g16 + i);
                } else if (i < n16) {                                               // the last piece of the input, byte by byte
                    const long long off = a0 + 16ll * i;
                    unsigned long long lo = 0ull, hi = 0ull;
                    for (int b = 0; b < 16 && off + b < total_bytes; ++b) {
                        const unsigned long long x = __ldg(raw + off + b);
                        if (b < 8) lo |= x << (8 * b);
                        else hi |= x << (8 * (b - 8));
                    }
                    v = make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
                }
                if (i < n16 + 2) s16[i] = v;
                const uint32_t nz = __ballot_sync(0xffffffffu, (v.x | v.y | v.z | v.w) != 0u);
                if (lane == 0 && i < n16 + 2) s_nz[i >> 5] = nz;
            }
        }
        __syncthreads();
        if (tid < nr) {
            const int nbytes = (S + 7) / 8;
            const unsigned int b0 = mis + (unsigned int)tid * (unsigned int)pitch_in;
            unsigned int b1 = b0 + (unsigned int)nbytes;
            uint32_t any = 0;
            if (S & 7) {                                          // pad bits of the last byte do not count (MSB-first)
                b1 -= 1;
                any = (smem_le32(s32, b1) & 0xffu) & (0xff00u >> (S & 7));
            }
            if (!any && b0 < b1) {
                const unsigned int m0 = min(b1, (b0 + 15u) & ~15u);       // [b0, m0): bytes shared with the row before
                const unsigned int m1 = max(m0, b1 & ~15u);               // [m1, b1): bytes shared with the row after
                unsigned int pa = m0 >> 4;
                const unsigned int pb = m1 >> 4;                          // pieces [pa, pb) lie wholly inside the row
                while (pa < pb && !any) {
                    const unsigned int wi = pa >> 5, lo = pa & 31u;
                    const unsigned int n = min(32u - lo, pb - pa);
                    const uint32_t m = (n == 32u ? 0xffffffffu : ((1u << n) - 1u)) << lo;
                    any = s_nz[wi] & m;
                    pa += n;
                }
                if (!any) any = smem_bytes_any(s32, b0, m0) | smem_bytes_any(s32, m1, b1);
            }
            s_flag[tid] = any ? 1 : 0;
        }
    }
    __syncthreads();
    const unsigned int keep = (tid < nr && s_flag[tid]) ? 1u : 0u;
    unsigned int total;
    const unsigned int ex = block_exclusive_scan(keep, &total);
    if (keep) s_src[ex] = (unsigned short)tid;
    if (FLAVOUR == 3) {
        // decoupled look-back by the whole CTA: warp w reads the 32 tile states at distance 32w .. 32w+31, so one round
        // trip covers 256 tiles (kThreads / 32 windows).  A window's result = the values from the nearest tile down to its
        // first inclusive prefix (if it holds one); the windows are then combined nearest first.
        if (tid == 0) st_relaxed_u64(state + 1 + tile, (tile == 0 ? kTilePrefix : kTileAgg) | total);
        unsigned long long excl = 0;
        long long j = tile - 1;
        bool done = tile == 0;
        while (!done) {
            const long long idx = j - 32ll * warp - lane;
            unsigned long long v;
            unsigned int pm, zm, need;
            int first_p;
            do {
                v = idx >= 0 ? ld_relaxed_u64(state + 1 + idx) : kTilePrefix;
                pm = __ballot_sync(0xffffffffu, (v >> 62) == 2ull);
                zm = __ballot_sync(0xffffffffu, (v >> 62) == 0ull);
                first_p = pm ? __ffs(pm) - 1 : 32;
                need = first_p >= 31 ? 0xffffffffu : ((2u << first_p) - 1u);
            } while (zm & need);
            unsigned long long c = lane <= first_p ? (v & kTileValue) : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (lane == 0) { s_lb_sum[warp] = c; s_lb_has[warp] = first_p < 32 ? 1u : 0u; }
            __syncthreads();
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) {
                if (done) break;
                excl += s_lb_sum[w];
                done = s_lb_has[w] != 0u;
            }
            __syncthreads();                                      // s_lb_* are rewritten by the next round
            j -= kThreads;
        }
        if (tid == 0) {
            if (tile != 0) st_relaxed_u64(state + 1 + tile, kTilePrefix | (excl + total));
            s_excl = excl;
            if (tile == n_tiles - 1) *d_chunk_total = (long long)(excl + total);
        }
    } else if (warp == 0) {
        // decoupled look-back: publish this tile's count, then sum the tiles before it.  The status word carries its
        // own payload (no other memory is handed over through it), so relaxed accesses are enough.
        unsigned long long excl = 0;
        if (tile == 0) {
            if (lane == 0) st_relaxed_u64(state + 1, kTilePrefix | total);
        } else {
            if (lane == 0) st_relaxed_u64(state + 1 + tile, kTileAgg | total);
            long long j = tile - 1;
            if (FLAVOUR == 0) {
                while (true) {
                    const long long idx = j - lane;
                    unsigned long long v;
                    unsigned int pm, zm, need;
                    int first_p;
                    do {
                        v = idx >= 0 ? ld_relaxed_u64(state + 1 + idx) : kTilePrefix;
                        pm = __ballot_sync(0xffffffffu, (v >> 62) == 2ull);
                        zm = __ballot_sync(0xffffffffu, (v >> 62) == 0ull);
                        first_p = pm ? __ffs(pm) - 1 : 32;
                        need = first_p >= 31 ? 0xffffffffu : ((2u << first_p) - 1u);
                    } while (zm & need);
                    unsigned long long c = lane <= first_p ? (v & kTileValue) : 0ull;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                    excl += c;
                    if (first_p < 32) break;
                    j -= 32;
                }
            } else {
                // window of 32 * kLook tiles per round trip: lane l reads the tiles at distance l*kLook .. l*kLook+kLook-1
                // (measured: wider windows poll more and are slower -- 0.287 ms at 128 tiles against 0.261 ms at 32)
                constexpr int kLook = 1;
                while (true) {
                    unsigned long long v[kLook];
                    unsigned int pm, bad;
                    int first_lane, my_p;
                    do {
                        my_p = kLook;                                 // nearest prefix among this lane's tiles
                        int my_z = kLook;                             // nearest tile that has not published yet
#pragma unroll
                        for (int e = 0; e < kLook; ++e) {
                            const long long idx = j - ((long long)lane * kLook + e);
                            v[e] = idx >= 0 ? ld_relaxed_u64(state + 1 + idx) : kTilePrefix;
                        }
#pragma unroll
                        for (int e = kLook - 1; e >= 0; --e) {
                            if ((v[e] >> 62) == 2ull) my_p = e;
                            if ((v[e] >> 62) == 0ull) my_z = e;
                        }
                        pm = __ballot_sync(0xffffffffu, my_p < kLook);
                        first_lane = pm ? __ffs(pm) - 1 : 32;
                        // unpublished tile in front of the nearest prefix -> poll again
                        bad = __ballot_sync(0xffffffffu, lane < first_lane ? my_z < kLook : (lane == first_lane && my_z < my_p));
                    } while (bad);
                    unsigned long long c = 0ull;
#pragma unroll
                    for (int e = 0; e < kLook; ++e)
                        if (lane < first_lane || (lane == first_lane && e <= my_p)) c += v[e] & kTileValue;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                    excl += c;
                    if (first_lane < 32) break;
                    j -= 32 * kLook;
                }
            }
            if (lane == 0) st_relaxed_u64(state + 1 + tile, kTilePrefix | (excl + total));
        }
        if (lane == 0) {
            s_excl = excl;
            if (tile == n_tiles - 1) *d_chunk_total = (long long)(excl + total);
        }
    }
    __syncthreads();
    if (total == 0) return;
    const long long base = *d_nrows + (long long)s_excl;
    const int q4 = pitchW >> 2;
    uint4 *out16 = reinterpret_cast<uint4 *>(rows_out + base * pitchW);
    const int units = (int)total * q4;
    const float inv_q4 = 1.0f / (float)q4;
    for (int u = tid; u < units; u += kThreads) {
        int k, q;
        if (FLAVOUR == 0) {
            k = u / q4;
            q = u - k * q4;
        } else {                                                  // units < 2^20: the float quotient is off by one at most
            k = (int)__fmul_rz((float)u, inv_q4);
            q = u - k * q4;
            if (q < 0) { k -= 1; q += q4; }
            else if (q >= q4) { k += 1; q -= q4; }
        }
        const unsigned int o = mis + (unsigned int)s_src[k] * (unsigned int)pitch_in + 16u * q;
        const unsigned int a = o >> 2, sh = (o & 3u) * 8u;
        const uint32_t r0 = s32[a], r1 = s32[a + 1], r2 = s32[a + 2], r3 = s32[a + 3], r4 = s32[a + 4];
        uint32_t w[4] = {__funnelshift_r(r0, r1, sh), __funnelshift_r(r1, r2, sh), __funnelshift_r(r2, r3, sh),
                         __funnelshift_r(r3, r4, sh)};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int wi = q * 4 + e;
            uint32_t x = wi < nW ? msb_bytes_to_word(w[e]) : 0u;
            if (wi == nW - 1) x &= last_mask;
            w[e] = x;
        }
        out16[u] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    if (af_out && af_in)
        for (int k = tid; k < (int)total; k += kThreads) af_out[base + k] = af_in[row0 + s_src[k]];
}

// ------------------------------------------------------------------------------------------------
// dense hdf5 rows (bool bytes, or float32 GT*AF).  Every row is kept (the file was filtered when it was written),
// so there is nothing to compact: one thread packs 32 consecutive samples of one row into one output word,
// neighbouring threads take neighbouring words (a warp reads 1 KB / 4 KB of contiguous input).  The float flavour
// also recovers the row's AF (all nonzero entries of a GT*AF row are equal, select.py:222): order-preserving
// atomicMax on the float bits, converted to float64 by a second, tiny kernel.
// ------------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256) pack_dense_kernel(const void *__restrict__ raw, long long n_rows, int S, int pitchW,
                                                         const long long *d_nrows, uint32_t *__restrict__ rows_out,
                                                         unsigned int *__restrict__ af_bits)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long r = idx / pitchW;
    if (r >= n_rows) return;
    const int w = (int)(idx - r * pitchW);
    const int s0 = w * 32;
    uint32_t word = 0;
    if (s0 < S) {
        const int n = min(32, S - s0);
        if (KIND == RAW_DENSE_U8) {
            const uint8_t *src = (const uint8_t *)raw + r * (long long)S + s0;
            if (n == 32 && (((uintptr_t)src) & 15u) == 0) {
                const uint4 a = ld_stream_u128(reinterpret_cast<const uint4 *>(src));
                const uint4 b = ld_stream_u128(reinterpret_cast<const uint4 *>(src) + 1);
                const uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    // nonzero bytes -> one bit each (byte j of word k is sample 4k+j)
                    const uint32_t nz = __vcmpne4(v[k], 0u);                 // 0xff per nonzero byte
                    const uint32_t bits = ((nz & 0x01u)) | ((nz >> 7) & 0x02u) | ((nz >> 14) & 0x04u) | ((nz >> 21) & 0x08u);
                    word |= bits << (4 * k);
                }
            } else {
                for (int j = 0; j < n; ++j) word |= (__ldg(src + j) != 0 ? 1u : 0u) << j;
            }
        } else {
            const float *src = (const float *)raw + r * (long long)S + s0;
            float vmax = 0.0f;
            if (n == 32 && (((uintptr_t)src) & 15u) == 0) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint4 a = ld_stream_u128(reinterpret_cast<const uint4 *>(src) + k);
                    const float f[4] = {__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(a.w)};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (f[j] != 0.0f) { word |= 1u << (4 * k + j); vmax = fmaxf(vmax, f[j]); }
                    }
                }
            } else {
                for (int j = 0; j < n; ++j) {
                    const float f = __ldg(src + j);
                    if (f != 0.0f) { word |= 1u << j; vmax = fmaxf(vmax, f); }
                }
            }
            // positive floats compare like their bit patterns; negative / NaN entries make the AF check at finalize fail
            if (word && af_bits) atomicMax(af_bits + r, __float_as_uint(vmax));
        }
    }
    rows_out[(*d_nrows + r) * pitchW + w] = word;
}

// ------------------------------------------------------------------------------------------------
// hdf5 bool chunks decoded ON the GPU (opt-in, UTMOS_B200_H5_GPU_LZF=1): the host only reads the compressed chunks
// (a few KB each for sparse genotypes) and copies them over PCIe; one warp per chunk decodes the LZF stream straight
// into a BIT buffer in shared memory -- output byte p becomes bit p (byte != 0), so the 250 KB a dense chunk would
// occupy are never materialised -- and then writes the chunk's rows in the matrix layout.  Token logic (hostio.cu has
// the byte decoder it mirrors; the word-level scheme was modelled in NumPy against it):
//   literal run (<= 32 bytes): one byte per lane, the ballot is inserted at bit p;
//   back reference, offset >= length: each lane builds one destination word with a funnel shift over the source bits;
//   back reference, offset < length (a repeating pattern, the way runs are stored): nothing to do when the pattern is
//   all zero (the buffer starts zeroed) -- the common case for sparse data --, else bit by bit with i % offset.
// ------------------------------------------------------------------------------------------------
constexpr int kLzfCompSmem = 8 * 1024;          // compressed bytes of a chunk staged in shared memory when they fit

__device__ __forceinline__ uint32_t bits_get32(const uint32_t *w, int n_words, long long sbit)
{
    const long long i = sbit >> 5;                         // arithmetic shift: floor for negative positions
    const unsigned int sh = (unsigned int)(sbit & 31);
    const uint32_t lo = (i >= 0 && i < n_words) ? w[i] : 0u;
    const uint32_t hi = (i + 1 >= 0 && i + 1 < n_words) ? w[i + 1] : 0u;
    return __funnelshift_r(lo, hi, sh);
}

__global__ void __launch_bounds__(32) lzf_unpack_bool_kernel(const uint8_t *__restrict__ blob, const long long *__restrict__ off,
                                                             const int *__restrict__ len, const uint8_t *__restrict__ stored_raw,
                                                             int rows_per_chunk, long long rows_in_batch, int S, int pitchW,
                                                             long long row0, uint32_t *__restrict__ rows_out, int *bad)
{
    extern __shared__ __align__(16) uint32_t z_smem[];
    const int lane = threadIdx.x;
    const long long chunk = blockIdx.x;
    const long long chunk_bytes = (long long)rows_per_chunk * S;
    const int n_words = (int)((chunk_bytes + 31) / 32) + 2;
    uint32_t *s_bits = z_smem;
    uint8_t *s_comp = reinterpret_cast<uint8_t *>(z_smem + n_words);
    for (int i = lane; i < n_words; i += 32) s_bits[i] = 0u;
    const uint8_t *src = blob + off[chunk];
    const int n = len[chunk];
    const bool staged = n <= kLzfCompSmem;
    if (staged)
        for (int i = lane; i < n; i += 32) s_comp[i] = src[i];
    __syncwarp();
    auto byte_at = [&](int i) -> unsigned int { return staged ? s_comp[i] : (unsigned int)__ldg(src + i); };
    bool ok = true;
    if (stored_raw[chunk]) {
        // unfiltered chunk: the bytes themselves
        if (n != chunk_bytes) ok = false;
        for (long long p0 = 0; ok && p0 < chunk_bytes; p0 += 32) {
            const long long p = p0 + lane;
            const uint32_t m = __ballot_sync(0xffffffffu, p < chunk_bytes && __ldg(src + p) != 0);
            if (lane == 0) s_bits[p0 >> 5] = m;
        }
    } else {
        int ip = 0;
        long long p = 0;
        while (ok && ip < n) {
            const unsigned int ctrl = byte_at(ip++);
            if (ctrl < 32u) {
                const int ln = (int)ctrl + 1;
                if (ip + ln > n || p + ln > chunk_bytes) { ok = false; break; }
                const uint32_t m = __ballot_sync(0xffffffffu, lane < ln && byte_at(ip + lane) != 0u);
                if (lane == 0 && m) {
                    const int w = (int)(p >> 5), sh = (int)(p & 31);
                    s_bits[w] |= m << sh;
                    if (sh + ln > 32) s_bits[w + 1] |= m >> (32 - sh);
                }
                ip += ln;
                p += ln;
            } else {
                int ln = (int)(ctrl >> 5);
                if (ln == 7) {
                    if (ip >= n) { ok = false; break; }
                    ln += (int)byte_at(ip++);
                }
                if (ip >= n) { ok = false; break; }
                const long long offb = (long long)(((ctrl & 31u) << 8) | byte_at(ip++)) + 1;
                ln += 2;
                if (offb > p || p + ln > chunk_bytes) { ok = false; break; }
                const long long q = p - offb;
                if (offb >= ln) {
                    const int wd0 = (int)(p >> 5), wd1 = (int)((p + ln - 1) >> 5);
                    const int wd = wd0 + lane;
                    uint32_t v = 0u;
                    if (wd <= wd1) {
                        v = bits_get32(s_bits, n_words, (long long)wd * 32 - offb);
                        const int lo_d = (int)(max(p, (long long)wd * 32) - (long long)wd * 32);
                        const int hi_d = (int)(min(p + ln, (long long)wd * 32 + 32) - (long long)wd * 32);
                        const uint32_t m = (hi_d == 32 ? 0xffffffffu : ((1u << hi_d) - 1u)) & ~((1u << lo_d) - 1u);
                        v &= m;
                    }
                    __syncwarp();                              // every source word has been read
                    if (v) s_bits[wd] |= v;
                } else {
                    // pattern of offb (< 264) bits repeated: is there a set bit in it at all?
                    uint32_t any = 0u;
                    for (long long b0 = (q >> 5) * 32 + 32ll * lane; b0 < p; b0 += 32 * 32) {
                        uint32_t x = s_bits[b0 >> 5];
                        if (b0 < q) x &= ~((1u << (q - b0)) - 1u);                 // bits below q
                        if (b0 + 32 > p) x &= (1u << (p - b0)) - 1u;               // bits from p on (p - b0 < 32 here)
                        any |= x;
                    }
                    if (__any_sync(0xffffffffu, any != 0u)) {
                        bool set[9];
#pragma unroll
                        for (int t = 0; t < 9; ++t) {
                            const int i = lane + 32 * t;
                            const long long sb = q + (i % (int)offb);
                            set[t] = i < ln && ((s_bits[sb >> 5] >> (sb & 31)) & 1u);
                        }
                        __syncwarp();
#pragma unroll
                        for (int t = 0; t < 9; ++t) {
                            const long long d = p + lane + 32 * t;
                            if (set[t]) atomicOr(&s_bits[d >> 5], 1u << (d & 31));
                        }
                    }
                }
                p += ln;
            }
            __syncwarp();
        }
        if (ok && p != chunk_bytes) ok = false;
    }
    __syncwarp();
    if (!ok) {
        if (lane == 0) atomicExch(bad, 1);
        return;
    }
    // rows of this chunk in the matrix layout (rows past the end of the dataset are padding of the last chunk)
    const int nW = (S + 31) / 32;
    for (int r = 0; r < rows_per_chunk; ++r) {
        const long long row = chunk * rows_per_chunk + r;
        if (row >= rows_in_batch) break;
        uint32_t *dst = rows_out + (row0 + row) * pitchW;
        for (int w = lane; w < pitchW; w += 32) {
            uint32_t v = 0u;
            if (w < nW) {
                v = bits_get32(s_bits, n_words, (long long)r * S + 32ll * w);
                const int nb = S - w * 32;
                if (nb < 32) v &= (1u << nb) - 1u;
            }
            dst[w] = v;
        }
    }
}

__global__ void dense_af_kernel(const unsigned int *af_bits, long long n_rows, const long long *d_nrows, double *af_out)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_rows) af_out[*d_nrows + r] = (double)__uint_as_float(af_bits[r]);
}

__global__ void bump_rows_by_kernel(long long *d_nrows, long long n) { *d_nrows += n; }

__global__ void bump_rows_kernel(long long *d_nrows, const long long *d_chunk_total) { *d_nrows += *d_chunk_total; }

}  // namespace

int ingest_scratch_reserve(IngestScratch &sc, long long rows, cudaStream_t stream)
{
    if (rows <= sc.cap_rows) return UTMOS_OK;
    ingest_scratch_free(sc, stream);
    const long long nb = (rows + kRowsPerBlock - 1) / kRowsPerBlock;
    UT_CUDA(cudaMallocAsync(&sc.flags, (size_t)rows * 4, stream));      // bytes: row flags; dense float flavour: one uint32 per row
    UT_CUDA(cudaMallocAsync(&sc.block_counts, sizeof(unsigned int) * (size_t)nb, stream));
    UT_CUDA(cudaMallocAsync(&sc.block_offsets, sizeof(unsigned int) * (size_t)nb, stream));
    UT_CUDA(cudaMallocAsync(&sc.tile_state, sizeof(unsigned long long) * (size_t)(rows + 2), stream));
    sc.cap_rows = rows;
    return UTMOS_OK;
}

void ingest_scratch_free(IngestScratch &sc, cudaStream_t stream)
{
    if (sc.flags) cudaFreeAsync(sc.flags, stream);
    if (sc.block_counts) cudaFreeAsync(sc.block_counts, stream);
    if (sc.block_offsets) cudaFreeAsync(sc.block_offsets, stream);
    if (sc.tile_state) cudaFreeAsync(sc.tile_state, stream);
    sc = IngestScratch();
}

template <int KIND>
static int ingest_kind(cudaStream_t stream, IngestScratch &sc, const void *raw, long long n_rows, long long pitch_in,
                       const double *af_in, double *af_tmp, int S, int pitchW, uint32_t *rows_out, double *af_out,
                       long long *d_nrows, long long *d_chunk_total)
{
    const long long nb = (n_rows + kRowsPerBlock - 1) / kRowsPerBlock;
    flag_rows_kernel<KIND><<<(unsigned)nb, kThreads, 0, stream>>>(raw, n_rows, pitch_in, S, sc.flags,
                                                                  sc.block_counts, af_tmp);
    scan_blocks_kernel<<<1, 1024, 0, stream>>>(sc.block_counts, sc.block_offsets, nb, d_chunk_total);
    scatter_rows_kernel<KIND><<<(unsigned)nb, kThreads, 0, stream>>>(
        raw, n_rows, pitch_in, KIND == RAW_DENSE_F32 ? af_tmp : af_in, S, pitchW, sc.flags, sc.block_offsets, d_nrows,
        rows_out, af_out);
    bump_rows_kernel<<<1, 1, 0, stream>>>(d_nrows, d_chunk_total);
    UT_CUDA(cudaGetLastError());
    return UTMOS_OK;
}

// d_nrows points at two consecutive device int64: [0] rows stored so far, [1] scratch for the chunk total.
// af_in: per-row AF of the chunk (device) or null; for the float flavour it must point at a writable
// scratch of n_rows doubles (filled by k1).
int launch_ingest(cudaStream_t stream, IngestScratch &sc, int kind, const void *raw, long long n_rows,
                  long long pitch_in, const double *af_in, int S, int pitchW, uint32_t *rows_out, double *af_out,
                  long long *d_nrows, int *n_launch)
{
    if (n_rows <= 0) return UTMOS_OK;
    UT_TRY(ingest_scratch_reserve(sc, n_rows, stream));
    long long *d_total = d_nrows + 1;
    if (kind == RAW_PACKED_MSB && ((uintptr_t)raw & 15u) == 0 && (size_t)pitch_in + 64 <= kFastSmemRaw &&
        n_rows * pitch_in < (1ll << 46)) {
        static size_t tile_pref = 0;
        if (!tile_pref) {
            const char *env = getenv("UTMOS_B200_INGEST_TILE");   // bytes staged per CTA (A/B runs)
            const long long want = env ? atoll(env) : 0;
            tile_pref = want >= 1024 && (size_t)want <= kFastSmemRaw ? (size_t)want : kFastTileBytes;
        }
        const size_t tile_budget = std::max(tile_pref, (size_t)pitch_in + 64);
        const int R = (int)std::min<long long>(kFastMaxRows, (long long)((tile_budget - 64) / (size_t)pitch_in));
        const long long n_tiles = (n_rows + R - 1) / R;
        const size_t smem = ((size_t)R * (size_t)pitch_in + 31) / 16 * 16 + 48;
        static bool configured = false;
        static int flavour = 3;               // TMA bulk load + CTA-wide look-back (round 2: 0.51 of the measured HBM peak)
        if (!configured) {
            const int smem_max = (int)(kFastSmemRaw + 64);
            UT_CUDA(cudaFuncSetAttribute(ingest_packed_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
            UT_CUDA(cudaFuncSetAttribute(ingest_packed_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
            UT_CUDA(cudaFuncSetAttribute(ingest_packed_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
            UT_CUDA(cudaFuncSetAttribute(ingest_packed_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
            const char *env = getenv("UTMOS_B200_INGEST");       // A/B runs: 0 = first version, 1 = piece bitmap, 2 = batched loads
            if (env) flavour = atoi(env) == 0 ? 0 : atoi(env) == 2 ? 2 : atoi(env) == 1 ? 1 : 3;
            configured = true;
        }
        UT_CUDA(cudaMemsetAsync(sc.tile_state, 0, sizeof(unsigned long long) * (size_t)(n_tiles + 1), stream));
        auto kernel = flavour == 0 ? ingest_packed_kernel<0> : flavour == 2 ? ingest_packed_kernel<2>
                      : flavour == 3 ? ingest_packed_kernel<3> : ingest_packed_kernel<1>;
        kernel<<<(unsigned)n_tiles, kThreads, smem, stream>>>((const uint8_t *)raw, n_rows, pitch_in, af_in, S, pitchW, R, n_tiles,
                                                              sc.tile_state, d_nrows, d_total, rows_out, af_out);
        bump_rows_kernel<<<1, 1, 0, stream>>>(d_nrows, d_total);
        *n_launch += 2;
        UT_CUDA(cudaGetLastError());
        return UTMOS_OK;
    }
    if (kind == RAW_DENSE_U8 || kind == RAW_DENSE_F32) {
        const long long items = n_rows * pitchW;
        const unsigned grid = (unsigned)((items + 255) / 256);
        if (kind == RAW_DENSE_U8) {
            pack_dense_kernel<RAW_DENSE_U8><<<grid, 256, 0, stream>>>(raw, n_rows, S, pitchW, d_nrows, rows_out, nullptr);
            *n_launch += 2;
        } else {
            unsigned int *af_bits = reinterpret_cast<unsigned int *>(sc.flags);     // scratch: >= 4 bytes per row (see reserve)
            UT_CUDA(cudaMemsetAsync(af_bits, 0, (size_t)n_rows * 4, stream));
            pack_dense_kernel<RAW_DENSE_F32><<<grid, 256, 0, stream>>>(raw, n_rows, S, pitchW, d_nrows, rows_out, af_bits);
            dense_af_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, stream>>>(af_bits, n_rows, d_nrows, af_out);
            *n_launch += 3;
        }
        bump_rows_by_kernel<<<1, 1, 0, stream>>>(d_nrows, n_rows);
        UT_CUDA(cudaGetLastError());
        return UTMOS_OK;
    }
    *n_launch += 4;
    switch (kind) {
    case RAW_PACKED_MSB:
        return ingest_kind<RAW_PACKED_MSB>(stream, sc, raw, n_rows, pitch_in, af_in, nullptr, S, pitchW, rows_out,
                                           af_out, d_nrows, d_total);
    case RAW_DENSE_U8:
        return ingest_kind<RAW_DENSE_U8>(stream, sc, raw, n_rows, pitch_in, af_in, nullptr, S, pitchW, rows_out,
                                         af_out, d_nrows, d_total);
    case RAW_DENSE_F32:
        return ingest_kind<RAW_DENSE_F32>(stream, sc, raw, n_rows, pitch_in, nullptr, const_cast<double *>(af_in), S,
                                          pitchW, rows_out, af_out, d_nrows, d_total);
    default:
        set_error("launch_ingest: unknown raw kind");
        return UTMOS_E_ARG;
    }
}

// ------------------------------------------------------------------------------------------------
// ".jl v2" rows (the "both axis pack" the reference's README floats, README.md:59-63): a row is stored
// either as its np.packbits bytes (length == pitch) or, when that is shorter, as the list of its carriers' sample
// indices (little-endian uint16 / uint32).  One warp per row rebuilds the MSB-first packed bytes in shared memory and
// writes them to the raw staging buffer the ingest kernel reads, so PCIe carries ~10x fewer bytes for a sparse cohort.
// bad[0] counts rows whose encoding is malformed (length not a whole number of indices, index >= S).
// ------------------------------------------------------------------------------------------------
template <int IDX_BYTES>
__global__ void __launch_bounds__(256) unpack_rows2_kernel(const uint8_t *__restrict__ payload,
                                                           const unsigned long long *__restrict__ off, long long n_rows,
                                                           int S, long long pitch, uint8_t *__restrict__ raw_out, int *bad)
{
    extern __shared__ __align__(16) uint32_t u_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int row_words = (int)((pitch + 3) >> 2);
    uint32_t *s_row = u_smem + (size_t)wib * row_words;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const unsigned long long off0 = off[0];
    for (long long r = warp0; r < n_rows; r += nwarps) {
        const unsigned long long b = off[r] - off0, e = off[r + 1] - off0;
        const long long len = (long long)(e - b);
        const uint8_t *src = payload + b;
        uint8_t *out = raw_out + r * pitch;
        if (len == pitch) {
            for (long long k = lane; k < pitch; k += 32) out[k] = src[k];
            continue;
        }
        for (int k = lane; k < row_words; k += 32) s_row[k] = 0u;
        __syncwarp();
        bool ok = len >= 0 && len % IDX_BYTES == 0 && len < pitch;
        const long long n = ok ? len / IDX_BYTES : 0;
        for (long long i = lane; i < n; i += 32) {
            unsigned int idx = (unsigned int)src[i * IDX_BYTES] | ((unsigned int)src[i * IDX_BYTES + 1] << 8);
            if (IDX_BYTES == 4) idx |= ((unsigned int)src[i * 4 + 2] << 16) | ((unsigned int)src[i * 4 + 3] << 24);
            if (idx >= (unsigned int)S) { ok = false; continue; }
            atomicOr(s_row + (idx >> 5), 1u << (8u * ((idx >> 3) & 3u) + 7u - (idx & 7u)));       // MSB-first within the byte
        }
        __syncwarp();
        for (long long k = lane; k < pitch; k += 32) out[k] = (uint8_t)(s_row[k >> 2] >> (8 * (k & 3)));
        if (__any_sync(0xffffffffu, !ok) && lane == 0) atomicAdd(bad, 1);
        __syncwarp();
    }
}

int launch_unpack_rows2(cudaStream_t stream, const uint8_t *payload, const unsigned long long *off, long long n_rows, int S,
                        long long pitch, int idx_bytes, uint8_t *raw_out, int *bad, int *n_launch)
{
    if (n_rows <= 0) return UTMOS_OK;
    const size_t smem = (size_t)8 * (size_t)((pitch + 3) / 4) * 4;
    if (smem > 200 * 1024) { set_error("unpack_rows2: row too wide for the on-chip row buffer"); return UTMOS_E_ARG; }
    static size_t configured[2] = {0, 0};
    const int which = idx_bytes == 4 ? 1 : 0;
    if (smem > configured[which]) {
        if (which) UT_CUDA(cudaFuncSetAttribute(unpack_rows2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else UT_CUDA(cudaFuncSetAttribute(unpack_rows2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[which] = smem;
    }
    const unsigned grid = (unsigned)std::min<long long>((n_rows + 7) / 8, 148ll * 8);
    if (which) unpack_rows2_kernel<4><<<grid, 256, smem, stream>>>(payload, off, n_rows, S, pitch, raw_out, bad);
    else unpack_rows2_kernel<2><<<grid, 256, smem, stream>>>(payload, off, n_rows, S, pitch, raw_out, bad);
    *n_launch += 1;
    UT_CUDA(cudaGetLastError());
    return UTMOS_OK;
}

// hdf5 bool chunks decoded on the GPU.  blob/off/len/stored_raw: device copies of the compressed chunks of one batch;
// the rows go to rows_out[(row0 + chunk * rows_per_chunk + r) * pitchW] and *d_nrows grows by rows_in_batch.
// Returns UTMOS_E_ARG when a chunk does not fit the kernel's shared-memory bit buffer (caller falls back).
int launch_lzf_unpack_bool(cudaStream_t stream, const uint8_t *blob, const long long *off, const int *len,
                           const uint8_t *stored_raw, long long n_chunks, int rows_per_chunk, long long rows_in_batch,
                           int S, int pitchW, long long row0, uint32_t *rows_out, long long *d_nrows, int *bad,
                           int *n_launch)
{
    const long long chunk_bytes = (long long)rows_per_chunk * S;
    const size_t smem = ((size_t)((chunk_bytes + 31) / 32) + 2) * 4 + kLzfCompSmem;
    if (smem > 200 * 1024 || n_chunks > 0x7fffffffll) { set_error("lzf_unpack: chunk too large for the on-chip bit buffer"); return UTMOS_E_ARG; }
    static size_t configured = 0;
    if (smem > configured) {
        UT_CUDA(cudaFuncSetAttribute(lzf_unpack_bool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    lzf_unpack_bool_kernel<<<(unsigned)n_chunks, 32, smem, stream>>>(blob, off, len, stored_raw, rows_per_chunk, rows_in_batch,
                                                                     S, pitchW, row0, rows_out, bad);
    bump_rows_by_kernel<<<1, 1, 0, stream>>>(d_nrows, rows_in_batch);
    *n_launch += 2;
    UT_CUDA(cudaGetLastError());
    return UTMOS_OK;
}

}  // namespace utmos
