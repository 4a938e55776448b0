// ingest.cu -- K2: raw chunk (packed .jl rows, dense bool rows or dense float32 GT*AF rows, already in HBM)
//              -> order-preserving compaction of the informative rows into the variant-major bit matrix.
//
// Replaces utmos/select.py:275-280 (np.unpackbits(count=S) -> .any(axis=1) -> boolean row filter on GT and
// AF) without ever unpacking: rows stay bit packed, only the bit order changes from np.packbits' MSB-first
// bytes to LSB-first 32-bit words (sample s = word s>>5, bit s&31) and the pitch is padded to 16 bytes so
// every later kernel can use aligned 128-bit loads.
//
// Algorithmic bytes per chunk: read V*P (+8V AF), write V'*pitch (+8V').
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

__global__ void bump_rows_kernel(long long *d_nrows, const long long *d_chunk_total) { *d_nrows += *d_chunk_total; }

}  // namespace

int ingest_scratch_reserve(IngestScratch &sc, long long rows, cudaStream_t stream)
{
    if (rows <= sc.cap_rows) return UTMOS_OK;
    ingest_scratch_free(sc, stream);
    const long long nb = (rows + kRowsPerBlock - 1) / kRowsPerBlock;
    UT_CUDA(cudaMallocAsync(&sc.flags, (size_t)rows, stream));
    UT_CUDA(cudaMallocAsync(&sc.block_counts, sizeof(unsigned int) * (size_t)nb, stream));
    UT_CUDA(cudaMallocAsync(&sc.block_offsets, sizeof(unsigned int) * (size_t)nb, stream));
    sc.cap_rows = rows;
    return UTMOS_OK;
}

void ingest_scratch_free(IngestScratch &sc, cudaStream_t stream)
{
    if (sc.flags) cudaFreeAsync(sc.flags, stream);
    if (sc.block_counts) cudaFreeAsync(sc.block_counts, stream);
    if (sc.block_offsets) cudaFreeAsync(sc.block_offsets, stream);
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
    *n_launch += 4;
    long long *d_total = d_nrows + 1;
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

}  // namespace utmos
