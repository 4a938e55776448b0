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

#include "common.cuh"

namespace utmos {

namespace {

template <int PLOIDY>
__global__ void __launch_bounds__(256) gt_pack_af_kernel(const int8_t *__restrict__ gt, long long V, int S, int ploidy_rt,
                                                         uint8_t *__restrict__ packed, long long pitch,
                                                         double *__restrict__ af, unsigned long long *het_hom,
                                                         uint8_t *__restrict__ singleton)
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
                if (singleton) singleton[r] = (s_hist[1] == 1u || s_zero == 1u) ? 1 : 0;
                if (s_het) atomicAdd(het_hom + 0, (unsigned long long)s_het);
                if (s_hom) atomicAdd(het_hom + 1, (unsigned long long)s_hom);
            }
        }
        __syncthreads();
    }
}

}  // namespace

int launch_convert_gt(cudaStream_t stream, const int8_t *gt, long long V, int S, int ploidy, uint8_t *packed,
                      long long pitch_out, double *af, unsigned long long *het_hom, uint8_t *singleton, int *n_launch)
{
    if (V <= 0) return UTMOS_OK;
    const unsigned grid = (unsigned)(V < 148ll * 16 ? V : 148ll * 16);
    if (ploidy == 2)
        gt_pack_af_kernel<2><<<grid, 256, 0, stream>>>(gt, V, S, ploidy, packed, pitch_out, af, het_hom, singleton);
    else
        gt_pack_af_kernel<0><<<grid, 256, 0, stream>>>(gt, V, S, ploidy, packed, pitch_out, af, het_hom, singleton);
    *n_launch += 1;
    UT_CUDA(cudaGetLastError());
    return UTMOS_OK;
}

}  // namespace utmos
