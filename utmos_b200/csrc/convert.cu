// convert.cu -- K1 gt_pack_af: genotype tensor -> MSB-first bit-packed presence rows + per-variant AF +
// het/hom totals + singleton flags, in one pass over the int8 genotypes.
//
// Replaces utmos/convert.py:57-87:
//   is_het | is_hom_alt                      (:64-71)  presence; scikit-allel 1.3.5 semantics: het = every
//                                                      allele called and not all equal; hom_alt = every
//                                                      allele called, all equal, allele > 0
//   is_het.sum(), is_hom_alt.sum()           (:65,:69) grand totals -> stats
//   count_alleles().to_frequencies()[:,1:].max(axis=1) (:75)  AF = max alt-allele count / called alleles
//                                                      (one IEEE float64 divide; NaN when nothing is called)
//   ac.is_singleton(1) | ac.is_singleton(0)  (:58-60)  flags for --no-singleton (rows dropped by the host)
//   np.packbits(axis=1)                      (:85)     MSB-first bytes, zero padded
//
// Algorithmic bytes: read ploidy*V*S (int8), write V*ceil(S/8) + 9*V.
#include <math_constants.h>

#include <algorithm>

#include "common.cuh"

namespace utmos {

namespace {

template <int PLOIDY>
__global__ void __launch_bounds__(256) gt_pack_af_kernel(const int8_t *__restrict__ gt, long long V, int S, int ploidy_rt,
                                                         uint8_t *__restrict__ packed, long long pitch,
                                                         double *__restrict__ af, unsigned long long *het_hom,
                                                         uint8_t *__restrict__ singleton, int drop_single)
{
    __shared__ unsigned int s_hist[128];       // counts of alleles 1..127 (allele 0 is counted in registers)
    __shared__ unsigned int s_an, s_zero, s_het, s_hom;
    const int ploidy = PLOIDY > 0 ? PLOIDY : ploidy_rt;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int nW = (S + 31) / 32;
    for (long long r = blockIdx.x; r < V; r += gridDim.x) {
        if (threadIdx.x < 128) s_hist[threadIdx.x] = 0;
        if (threadIdx.x == 0) { s_an = 0; s_zero = 0; s_het = 0; s_hom = 0; }
        __syncthreads();
        const int8_t *row = gt + (size_t)r * (size_t)S * (size_t)ploidy;
        uint8_t *out = packed + r * pitch;
        unsigned int an = 0, zero = 0, het = 0, hom = 0;
        for (int w = warp; w < nW; w += nwarp) {
            const int s = w * 32 + lane;
            bool is_het = false, is_hom = false;
            if (s < S) {
                int g[PLOIDY > 0 ? PLOIDY : 8];
                if (PLOIDY == 2) {
                    const char2 v = *reinterpret_cast<const char2 *>(row + 2 * (size_t)s);
                    g[0] = v.x; g[1] = v.y;
                } else {
                    for (int k = 0; k < ploidy; ++k) g[k] = row[(size_t)s * ploidy + k];
                }
                bool called = true, equal = true;
                for (int k = 0; k < ploidy; ++k) {
                    if (g[k] < 0) called = false;
                    else {
                        an += 1;
                        if (g[k] == 0) zero += 1; else atomicAdd(&s_hist[g[k]], 1u);
                    }
                    if (g[k] != g[0]) equal = false;
                }
                is_het = called && !equal;
                is_hom = called && equal && g[0] > 0;
            }
            const uint32_t hw = __ballot_sync(0xffffffffu, is_het);
            const uint32_t mw = __ballot_sync(0xffffffffu, is_hom);
            if (lane == 0) { het += __popc(hw); hom += __popc(mw); }
            const uint32_t bytes = word_to_msb_bytes(hw | mw);
            if (lane < 4 && (long long)w * 4 + lane < pitch) out[w * 4 + lane] = (uint8_t)(bytes >> (8 * lane));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            an += __shfl_xor_sync(0xffffffffu, an, o);
            zero += __shfl_xor_sync(0xffffffffu, zero, o);
        }
        if (lane == 0) {
            atomicAdd(&s_an, an);
            atomicAdd(&s_zero, zero);
            if (het) atomicAdd(&s_het, het);
            if (hom) atomicAdd(&s_hom, hom);
        }
        __syncthreads();
        if (warp == 0) {
            unsigned int m = 0;
            for (int a = 1 + lane; a < 128; a += 32) m = max(m, s_hist[a]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
            if (lane == 0) {
                const unsigned int n = s_an;
                // max over alleles of count/an == (max count)/an: the divide is monotone in the numerator
                af[r] = n ? (double)m / (double)n : CUDART_NAN;
                const bool single = s_hist[1] == 1u || s_zero == 1u;
                if (singleton) singleton[r] = single ? 1 : 0;
                if (!(drop_single && single)) {            // --no-singleton: the row is dropped before the stats (convert.py:58-69)
                    if (s_het) atomicAdd(het_hom + 0, (unsigned long long)s_het);
                    if (s_hom) atomicAdd(het_hom + 1, (unsigned long long)s_hom);
                }
            }
        }
        __syncthreads();
    }
}

// One 32-bit word = two diploid samples (allele bytes g0,g1 | g0,g1).  SIMD-in-register byte compares; returns the two
// presence bits (bit 7: first sample, bit 23: second) and accumulates the per-row counters.
struct RowAcc {
    unsigned int an, zero, one, het, hom;
    bool rare;
};

__device__ __forceinline__ uint32_t gt_word(uint32_t v, RowAcc &a, unsigned int *hist)
{
    const uint32_t called = ~v & 0x80808080u;                 // bit 7 of every called allele byte
    a.an += __popc(called);
    a.zero += __popc(__vcmpeq4(v, 0u) & 0x01010101u);
    a.one += __popc(__vcmpeq4(v, 0x01010101u) & 0x01010101u);
    if (__vcmpgts4(v, 0x01010101u)) {                         // alleles >= 2: rare, per-warp histogram
        a.rare = true;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int g = (int)(int8_t)(v >> (8 * e));
            if (g >= 2) atomicAdd(&hist[g], 1u);
        }
    }
    const uint32_t both = called & __byte_perm(called, 0u, 0x2301);        // both alleles of a sample called
    const uint32_t eq = __vcmpeq4(v, __byte_perm(v, 0u, 0x2301));          // 0xff.. where allele0 == allele1
    const uint32_t pos = __vcmpgts4(v, 0u);                                // 0xff where allele > 0
    const uint32_t het_b = both & ~eq & 0x00800080u;
    const uint32_t hom_b = both & eq & pos & 0x00800080u;
    a.het += __popc(het_b);
    a.hom += __popc(hom_b);
    return het_b | hom_b;
}

// 16 genotype bytes = 8 diploid samples -> one MSB-first presence byte (sample j of the piece = bit 7 - j).
// Fast path: every allele of the piece is 0 or 1 (no missing call, no multi-allelic call) -- then presence = a0 | a1,
// hom-alt = a0 & a1, and the allele counts follow from the two popcounts (ones = present + hom, called = 16):
// ~11 integer instructions per 4 genotype bytes instead of ~33 for the byte-compare path (the video SIMD intrinsics
// are emulated on sm_100a).  `valid` = samples of the piece that exist (8 except at the ragged end of a row).
__device__ __forceinline__ uint32_t gt_piece(const uint4 &q, int valid, RowAcc &a, unsigned int *hist)
{
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    if (valid >= 8 && ((q.x | q.y | q.z | q.w) & 0xfefefefeu) == 0u) {
        uint32_t byte = 0u, homm = 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t v = w[k], sh = v >> 8;
            const uint32_t t = (v | sh) & 0x00010001u;          // bit 0: sample 2k present, bit 16: sample 2k+1 present
            byte |= (t << (7 - 2 * k)) | (t >> (10 + 2 * k));
            homm |= (v & sh & 0x00010001u) << k;
        }
        byte &= 0xffu;
        const unsigned int present = __popc(byte), hom = __popc(homm);
        a.an += 16u;
        a.hom += hom;
        a.het += present - hom;
        a.one += present + hom;
        a.zero += 16u - present - hom;
        return byte;
    }
    uint32_t byte = 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t v = w[k];
        if (2 * k >= valid) v = 0xffffffffu;                     // past the row (last group): missing
        else if (2 * k + 1 >= valid) v |= 0xffff0000u;
        const uint32_t pr = gt_word(v, a, hist);
        byte |= ((pr >> 7) & 1u) << (7 - 2 * k);
        byte |= ((pr >> 23) & 1u) << (6 - 2 * k);
    }
    return byte;
}

// warp-level end of a row: reduce the counters, AF = max alt-allele count / called alleles, singleton flag
__device__ __forceinline__ void gt_row_finish(RowAcc acc, unsigned int *hist, int lane, long long r, double *af,
                                              uint8_t *singleton, int drop_single, unsigned long long &het_tot,
                                              unsigned long long &hom_tot)
{
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) {
        acc.an += __shfl_xor_sync(0xffffffffu, acc.an, sh);
        acc.zero += __shfl_xor_sync(0xffffffffu, acc.zero, sh);
        acc.one += __shfl_xor_sync(0xffffffffu, acc.one, sh);
        acc.het += __shfl_xor_sync(0xffffffffu, acc.het, sh);
        acc.hom += __shfl_xor_sync(0xffffffffu, acc.hom, sh);
    }
    unsigned int m = acc.one;
    if (__any_sync(0xffffffffu, acc.rare)) {
        __syncwarp();
        for (int a = 2 + lane; a < 128; a += 32) { m = max(m, hist[a]); hist[a] = 0; }
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, sh));
        __syncwarp();
    }
    if (lane == 0) {
        // max over alleles of count/an == (max count)/an: the divide is monotone in the numerator
        af[r] = acc.an ? (double)m / (double)acc.an : CUDART_NAN;
        const bool single = acc.one == 1u || acc.zero == 1u;
        if (singleton) singleton[r] = single ? 1 : 0;
        if (!(drop_single && single)) {                // --no-singleton: the row is dropped before the stats (convert.py:58-69)
            het_tot += acc.het;
            hom_tot += acc.hom;
        }
    }
}

// Rows that start on 16-byte boundaries (2S % 16 == 0, e.g. S = 2504): no staging at all -- one warp per row, a lane
// loads 16 genotype bytes (8 samples) per turn straight from global memory with a streaming 128-bit load and emits
// one output byte; nothing but registers between the load and the store, so many warps keep loads in flight.
__global__ void __launch_bounds__(256) gt_pack_af_direct_kernel(const int8_t *__restrict__ gt, long long V, int S,
                                                                uint8_t *__restrict__ packed, long long pitch,
                                                                double *__restrict__ af, unsigned long long *het_hom,
                                                                uint8_t *__restrict__ singleton, int drop_single)
{
    __shared__ unsigned int s_hist[8][128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = lane; i < 128; i += 32) s_hist[warp][i] = 0;
    __syncwarp();
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int groups = (S + 7) >> 3;                       // 16-byte pieces per row == output bytes per row
    unsigned long long het_tot = 0, hom_tot = 0;
    for (long long r = warp0; r < V; r += nwarps) {
        const uint4 *row = reinterpret_cast<const uint4 *>(gt + r * 2ll * S);
        uint8_t *out = packed + r * pitch;
        RowAcc acc = {0u, 0u, 0u, 0u, 0u, false};
#pragma unroll 2
        for (int g = lane; g < groups; g += 32) {
            const uint4 q = ld_stream_u128(row + g);
            out[g] = (uint8_t)gt_piece(q, S - g * 8, acc, s_hist[warp]);
        }
        gt_row_finish(acc, s_hist[warp], lane, r, af, singleton, drop_single, het_tot, hom_tot);
    }
    if (lane == 0) {
        if (het_tot) atomicAdd(het_hom + 0, het_tot);
        if (hom_tot) atomicAdd(het_hom + 1, hom_tot);
    }
}

// ------------------------------------------------------------------------------------------------
// diploid fast path.  A CTA stages a tile of consecutive rows (one contiguous byte range of R * 2S genotype
// bytes) in shared memory with aligned 128-bit streaming loads, then one warp per row: a lane takes 8 samples
// (16 genotype bytes) per turn and emits exactly one MSB-first output byte, so a warp writes 32 consecutive
// bytes.  Allele 0 / allele 1 / called counts stay in registers and are shuffled down once per row; the rare
// alleles >= 2 go to a per-warp shared histogram.
// ------------------------------------------------------------------------------------------------
constexpr int kCvtWarps = 8;
constexpr size_t kCvtSmemRaw = 96 * 1024;        // largest row (2S bytes) the fast path stages
constexpr size_t kCvtTileBytes = 40 * 1024;      // preferred tile: 5 CTAs per SM overlap their load and compute phases

__device__ __forceinline__ uint32_t cvt_le32(const uint32_t *s32, unsigned int o)
{
    const unsigned int a = o >> 2;
    return __funnelshift_r(s32[a], s32[a + 1], (o & 3u) * 8u);
}

__global__ void __launch_bounds__(kCvtWarps * 32) gt_pack_af_tile_kernel(const int8_t *__restrict__ gt, long long V, int S,
                                                                         int rows_per_tile, uint8_t *__restrict__ packed,
                                                                         long long pitch, double *__restrict__ af,
                                                                         unsigned long long *het_hom,
                                                                         uint8_t *__restrict__ singleton, int drop_single)
{
    extern __shared__ __align__(16) uint8_t c_smem[];
    __shared__ unsigned int s_hist[kCvtWarps][128];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long row_bytes = 2ll * S;
    const long long total_bytes = V * row_bytes;
    for (int i = lane; i < 128; i += 32) s_hist[warp][i] = 0;
    unsigned long long het_tot = 0, hom_tot = 0;
    for (long long row0 = (long long)blockIdx.x * rows_per_tile; row0 < V; row0 += (long long)gridDim.x * rows_per_tile) {
        const int nr = (int)min((long long)rows_per_tile, V - row0);
        const long long start = row0 * row_bytes, end = start + nr * row_bytes;
        const long long a0 = start & ~15ll;
        const unsigned int mis = (unsigned int)(start - a0);
        const int n16 = (int)((end - a0 + 15) >> 4);
        uint4 *s16 = reinterpret_cast<uint4 *>(c_smem);
        __syncthreads();                                   // previous tile fully consumed
        for (int i = tid; i < n16 + 2; i += kCvtWarps * 32) {
            const long long off = a0 + 16ll * i;
            uint4 v = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);     // -1 = missing allele
            if (i < n16) {
                if (off + 16 <= total_bytes) {
                    v = ld_stream_u128(reinterpret_cast<const uint4 *>(gt + off));
                } else {
                    uint32_t w[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
                    for (int b = 0; b < 16 && off + b < total_bytes; ++b) {
                        w[b >> 2] &= ~(0xffu << (8 * (b & 3)));
                        w[b >> 2] |= (uint32_t)(uint8_t)__ldg(gt + off + b) << (8 * (b & 3));
                    }
                    v = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            s16[i] = v;
        }
        __syncthreads();
        const uint32_t *s32 = reinterpret_cast<const uint32_t *>(c_smem);
        for (int i = warp; i < nr; i += kCvtWarps) {
            const long long r = row0 + i;
            const unsigned int o = mis + (unsigned int)(i * row_bytes);
            uint8_t *out = packed + r * pitch;
            RowAcc acc = {0u, 0u, 0u, 0u, 0u, false};
            for (int s0 = lane * 8; s0 < S; s0 += 256) {
                // word k = samples s0+2k (bytes 0,1) and s0+2k+1 (bytes 2,3); five aligned words + funnel shifts
                const unsigned int ob = o + 2u * s0, wa = ob >> 2, fs = (ob & 3u) * 8u;
                const uint32_t r0 = s32[wa], r1 = s32[wa + 1], r2 = s32[wa + 2], r3 = s32[wa + 3], r4 = s32[wa + 4];
                const uint4 q = make_uint4(__funnelshift_r(r0, r1, fs), __funnelshift_r(r1, r2, fs), __funnelshift_r(r2, r3, fs),
                                           __funnelshift_r(r3, r4, fs));
                out[s0 >> 3] = (uint8_t)gt_piece(q, S - s0, acc, s_hist[warp]);
            }
            gt_row_finish(acc, s_hist[warp], lane, r, af, singleton, drop_single, het_tot, hom_tot);
        }
    }
    if (lane == 0) {
        if (het_tot) atomicAdd(het_hom + 0, het_tot);
        if (hom_tot) atomicAdd(het_hom + 1, hom_tot);
    }
}

}  // namespace

int launch_convert_gt(cudaStream_t stream, const int8_t *gt, long long V, int S, int ploidy, uint8_t *packed,
                      long long pitch_out, double *af, unsigned long long *het_hom, uint8_t *singleton, int drop_single,
                      int *n_launch)
{
    if (V <= 0) return UTMOS_OK;
    if (ploidy == 2 && ((uintptr_t)gt & 15u) == 0 && (2ll * S) % 16 == 0 && pitch_out == (S + 7) / 8) {
        const long long warps = V;
        const unsigned grid = (unsigned)std::min<long long>((warps + 7) / 8, 148ll * 16);
        gt_pack_af_direct_kernel<<<grid, 256, 0, stream>>>(gt, V, S, packed, pitch_out, af, het_hom, singleton, drop_single);
        *n_launch += 1;
        UT_CUDA(cudaGetLastError());
        return UTMOS_OK;
    }
    if (ploidy == 2 && ((uintptr_t)gt & 15u) == 0 && 2ull * (size_t)S + 64 <= kCvtSmemRaw && pitch_out == (S + 7) / 8) {
        const size_t tile_budget = std::max(kCvtTileBytes, 2 * (size_t)S + 64);
        int R = (int)std::min<size_t>(64, (tile_budget - 64) / (2 * (size_t)S));
        if (R > kCvtWarps) R = R / kCvtWarps * kCvtWarps;           // whole rounds of one row per warp
        const size_t smem = ((size_t)R * 2 * (size_t)S + 31) / 16 * 16 + 48;
        static bool configured = false;
        if (!configured) {
            UT_CUDA(cudaFuncSetAttribute(gt_pack_af_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(kCvtSmemRaw + 64)));
            configured = true;
        }
        const long long tiles = (V + R - 1) / R;
        const unsigned grid = (unsigned)std::min<long long>(tiles, 148ll * 8);
        gt_pack_af_tile_kernel<<<grid, kCvtWarps * 32, smem, stream>>>(gt, V, S, R, packed, pitch_out, af, het_hom, singleton, drop_single);
        *n_launch += 1;
        UT_CUDA(cudaGetLastError());
        return UTMOS_OK;
    }
    const unsigned grid = (unsigned)(V < 148ll * 16 ? V : 148ll * 16);
    if (ploidy == 2)
        gt_pack_af_kernel<2><<<grid, 256, 0, stream>>>(gt, V, S, ploidy, packed, pitch_out, af, het_hom, singleton, drop_single);
    else
        gt_pack_af_kernel<0><<<grid, 256, 0, stream>>>(gt, V, S, ploidy, packed, pitch_out, af, het_hom, singleton, drop_single);
    *n_launch += 1;
    UT_CUDA(cudaGetLastError());
    return UTMOS_OK;
}

}  // namespace utmos
