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
#include <stdlib.h>

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

// Rows that do NOT start on 16-byte boundaries (2S % 16 != 0, the general case): the same one-warp-per-row scheme without
// any staging.  A lane reads the two aligned 16-byte pieces its 16 genotype bytes straddle (plain read-only loads: the
// second piece is the neighbouring lane's first, served by L1) and realigns them in registers -- a warp-uniform word
// select on (offset >> 2) plus four funnel shifts by (offset & 3) bytes.  Replaces the shared-memory tile kernel for
// such rows (100,000 x 2,500: 0.205 ms against 0.276 ms for the tile kernel it replaced in round 2).
__device__ __forceinline__ uint4 realign16(const uint4 &a, const uint4 &b, unsigned int mis)
{
    uint32_t v0, v1, v2, v3, v4;
    switch (mis >> 2) {                                  // warp-uniform: the offset is a property of the row
    case 0: v0 = a.x; v1 = a.y; v2 = a.z; v3 = a.w; v4 = b.x; break;
    case 1: v0 = a.y; v1 = a.z; v2 = a.w; v3 = b.x; v4 = b.y; break;
    case 2: v0 = a.z; v1 = a.w; v2 = b.x; v3 = b.y; v4 = b.z; break;
    default: v0 = a.w; v1 = b.x; v2 = b.y; v3 = b.z; v4 = b.w; break;
    }
    const unsigned int sh = (mis & 3u) * 8u;
    return make_uint4(__funnelshift_r(v0, v1, sh), __funnelshift_r(v1, v2, sh), __funnelshift_r(v2, v3, sh),
                      __funnelshift_r(v3, v4, sh));
}

__global__ void __launch_bounds__(256) gt_pack_af_unaligned_kernel(const int8_t *__restrict__ gt, long long V, int S,
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
    const int groups = (S + 7) >> 3;                       // output bytes per row
    const long long row_bytes = 2ll * S;
    const long long n16_total = (V * row_bytes) >> 4;      // aligned pieces that lie wholly inside the tensor
    const uint4 *base16 = reinterpret_cast<const uint4 *>(gt);
    const uint4 missing = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    unsigned long long het_tot = 0, hom_tot = 0;
    for (long long r = warp0; r < V; r += nwarps) {
        const long long start = r * row_bytes;
        const long long p0 = start >> 4;
        const unsigned int mis = (unsigned int)(start & 15);
        uint8_t *out = packed + r * pitch;
        RowAcc acc = {0u, 0u, 0u, 0u, 0u, false};
#pragma unroll 2
        for (int g = lane; g < groups; g += 32) {
            const long long pa = p0 + g;
            uint4 a = missing, b = missing;
            if (pa < n16_total) a = __ldg(base16 + pa);
            else {                                          // the ragged last piece of the tensor, byte by byte
                uint32_t w[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
                for (int k = 0; k < 16 && pa * 16 + k < V * row_bytes; ++k) {
                    w[k >> 2] &= ~(0xffu << (8 * (k & 3)));
                    w[k >> 2] |= (uint32_t)(uint8_t)__ldg(gt + pa * 16 + k) << (8 * (k & 3));
                }
                a = make_uint4(w[0], w[1], w[2], w[3]);
            }
            if (mis) {
                if (pa + 1 < n16_total) b = __ldg(base16 + pa + 1);
                else if ((pa + 1) * 16 < V * row_bytes) {
                    uint32_t w[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
                    for (int k = 0; k < 16 && (pa + 1) * 16 + k < V * row_bytes; ++k) {
                        w[k >> 2] &= ~(0xffu << (8 * (k & 3)));
                        w[k >> 2] |= (uint32_t)(uint8_t)__ldg(gt + (pa + 1) * 16 + k) << (8 * (k & 3));
                    }
                    b = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            const uint4 q = mis ? realign16(a, b, mis) : a;
            out[g] = (uint8_t)gt_piece(q, S - g * 8, acc, s_hist[warp]);
        }
        gt_row_finish(acc, s_hist[warp], lane, r, af, singleton, drop_single, het_tot, hom_tot);
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
    if (ploidy == 2 && ((uintptr_t)gt & 15u) == 0 && pitch_out == (S + 7) / 8) {
        const unsigned grid = (unsigned)std::min<long long>((V + 7) / 8, 148ll * 16);
        gt_pack_af_unaligned_kernel<<<grid, 256, 0, stream>>>(gt, V, S, packed, pitch_out, af, het_hom, singleton, drop_single);
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
