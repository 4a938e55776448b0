// synth.cu -- deterministic synthetic cohorts generated directly in HBM (benchmark / test input only).
//
// allele count k_v ~ truncated power law, AF_v = k_v/(2S), carrier probability
// p_v = 1-(1-AF_v)^2, bit(v,s) ~ Bernoulli(p_v) from a counter-based hash of (seed, v, s), at least one
// carrier per row.  Both distributions enter through host-built integer tables (cdf_thr, p_thr), so the
// NumPy mirror in utmos_b200/synth.py reproduces every bit exactly.  Output is the .jl layout
// (MSB-first bytes, pitch ceil(S/8)) so it feeds utmos_append_packed[_device] like real data.
#include "common.cuh"

namespace utmos {
namespace {

__host__ __device__ __forceinline__ unsigned long long mix64(unsigned long long z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__host__ __device__ __forceinline__ unsigned long long cell_hash(unsigned long long seed, unsigned long long v,
                                                                  unsigned long long s)
{
    return mix64(seed + v * 0x9E3779B97F4A7C15ull + (s + 1ull) * 0xD1B54A32D192ED03ull);
}

// one warp per row
__global__ void __launch_bounds__(256) synth_rows_kernel(unsigned long long seed, long long row0, long long n_rows,
                                                         int S, const unsigned long long *__restrict__ cdf_thr,
                                                         const unsigned int *__restrict__ p_thr, int kmax,
                                                         uint8_t *__restrict__ out, double *__restrict__ af_out)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int nW = (S + 31) / 32;
    const long long pitch = (S + 7) / 8;
    for (long long i = warp; i < n_rows; i += nwarps) {
        const unsigned long long v = (unsigned long long)(row0 + i);
        // allele count: smallest k in [1, kmax] with u < cdf_thr[k]  (cdf_thr[kmax] = 2^64-1)
        const unsigned long long u = cell_hash(seed, v, 0xFFFFFFFFull);
        int lo = 1, hi = kmax;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (u < cdf_thr[mid]) hi = mid; else lo = mid + 1;
        }
        const int k = lo;
        const unsigned int thr = p_thr[k];
        uint8_t *row = out + i * pitch;
        uint32_t any = 0;
        for (int w0 = 0; w0 < nW; w0 += 32) {
            const int w = w0 + lane;
            uint32_t word = 0;
            if (w < nW) {
                for (int j = 0; j < 32; ++j) {
                    const int s = w * 32 + j;
                    if (s < S && (unsigned int)(cell_hash(seed, v, (unsigned long long)s) >> 32) < thr) word |= 1u << j;
                }
            }
            any |= __ballot_sync(0xffffffffu, word != 0);
            if (w < nW) {
                const uint32_t bytes = word_to_msb_bytes(word);
                for (int b = 0; b < 4; ++b)
                    if ((long long)w * 4 + b < pitch) row[w * 4 + b] = (uint8_t)(bytes >> (8 * b));
            }
        }
        if (!any && lane == 0) {
            const int s = (int)(cell_hash(seed, v, 0xFFFFFFFEull) % (unsigned long long)S);
            row[s >> 3] |= (uint8_t)(0x80u >> (s & 7));
        }
        if (lane == 0) af_out[i] = (double)k / (2.0 * (double)S);
    }
}

}  // namespace
}  // namespace utmos

using namespace utmos;

extern "C" {

int utmos_device_alloc(int device, void **ptr_out, int64_t bytes)
{
    if (!ptr_out || bytes < 0) { set_error("device_alloc: bad arguments"); return UTMOS_E_ARG; }
    UT_CUDA(cudaSetDevice(device));
    UT_CUDA(cudaMalloc(ptr_out, (size_t)(bytes > 0 ? bytes : 16)));
    return UTMOS_OK;
}

int utmos_device_free(int device, void *ptr)
{
    UT_CUDA(cudaSetDevice(device));
    if (ptr) UT_CUDA(cudaFree(ptr));
    return UTMOS_OK;
}

int utmos_device_to_host(int device, void *dst, const void *d_src, int64_t bytes)
{
    UT_CUDA(cudaSetDevice(device));
    UT_CUDA(cudaMemcpy(dst, d_src, (size_t)bytes, cudaMemcpyDeviceToHost));
    return UTMOS_OK;
}

// Fill d_rows (n_rows x ceil(S/8) bytes) and d_af (n_rows doubles), both DEVICE buffers, with rows
// row0 .. row0+n_rows-1 of the cohort identified by `seed`.  cdf_thr[kmax+1] / p_thr[kmax+1] are HOST tables.
int utmos_synth_packed_device(int device, uint64_t seed, int64_t row0, int64_t n_rows, int64_t n_samples,
                              const uint64_t *cdf_thr, const uint32_t *p_thr, int64_t kmax, uint8_t *d_rows,
                              double *d_af)
{
    if (n_rows < 0 || n_samples <= 0 || kmax < 1 || !cdf_thr || !p_thr) { set_error("synth: bad arguments"); return UTMOS_E_ARG; }
    if (n_rows == 0) return UTMOS_OK;
    UT_CUDA(cudaSetDevice(device));
    unsigned long long *d_cdf = nullptr;
    unsigned int *d_p = nullptr;
    UT_CUDA(cudaMalloc(&d_cdf, (size_t)(kmax + 1) * 8));
    cudaError_t e = cudaMalloc(&d_p, (size_t)(kmax + 1) * 4);
    if (e == cudaSuccess) e = cudaMemcpy(d_cdf, cdf_thr, (size_t)(kmax + 1) * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_p, p_thr, (size_t)(kmax + 1) * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        long long blocks = (n_rows + 7) / 8;
        if (blocks > 148 * 32) blocks = 148 * 32;
        synth_rows_kernel<<<(unsigned)blocks, 256>>>((unsigned long long)seed, row0, n_rows, (int)n_samples, d_cdf, d_p,
                                                     (int)kmax, d_rows, d_af);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
    }
    cudaFree(d_cdf);
    if (d_p) cudaFree(d_p);
    UT_CUDA(e);
    return UTMOS_OK;
}

}  // extern "C"
