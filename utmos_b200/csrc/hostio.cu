// hostio.cu -- host-side byte codecs used by the hdf5 chunk streamer (no device code).
//
// hdf5 filter 32000 ("lzf", the codec h5py applies for compression="lzf", utmos/select.py:208-238) stores each
// chunk as one liblzf block.  Format restated from the published liblzf stream description:
//   ctrl < 32           : literal run of ctrl+1 bytes
//   ctrl >= 32          : back reference, len = ctrl>>5 (if 7: + next byte), then +2;
//                         offset = ((ctrl & 31) << 8 | next byte) + 1 bytes behind the write position
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "common.cuh"

extern "C" {

// Returns the decoded length, or -1 on a malformed stream / output overflow.
int64_t utmos_lzf_decompress(const uint8_t *src, int64_t src_len, uint8_t *dst, int64_t dst_cap)
{
    int64_t ip = 0, op = 0;
    while (ip < src_len) {
        const unsigned ctrl = src[ip++];
        if (ctrl < 32) {
            const int64_t n = (int64_t)ctrl + 1;
            if (ip + n > src_len || op + n > dst_cap) return -1;
            memcpy(dst + op, src + ip, (size_t)n);
            ip += n;
            op += n;
        } else {
            int64_t len = ctrl >> 5;
            if (len == 7) {
                if (ip >= src_len) return -1;
                len += src[ip++];
            }
            if (ip >= src_len) return -1;
            const int64_t off = (((int64_t)(ctrl & 31)) << 8 | src[ip++]) + 1;
            len += 2;
            if (off > op || op + len > dst_cap) return -1;
            const uint8_t *ref = dst + op - off;
            uint8_t *out = dst + op;
            if (off >= len) memcpy(out, ref, (size_t)len);
            else for (int64_t i = 0; i < len; ++i) out[i] = ref[i];      // overlapping run
            op += len;
        }
    }
    return op;
}

// Greedy hash-chain-free LZF encoder.  Returns the encoded length, or 0 when the result would not fit in
// dst_cap (the caller then stores the chunk raw and sets filter_mask bit 0, as the hdf5 lzf filter does).
int64_t utmos_lzf_compress(const uint8_t *src, int64_t n, uint8_t *dst, int64_t dst_cap)
{
    enum { HLOG = 16, HSIZE = 1 << HLOG, MAX_OFF = 1 << 13, MAX_REF = (1 << 8) + (1 << 3), MAX_LIT = 32 };
    static thread_local int64_t htab[HSIZE];
    for (int i = 0; i < HSIZE; ++i) htab[i] = -1;
    int64_t ip = 0, op = 0, lit_start = 0;
    auto flush_literals = [&](int64_t end) -> bool {
        int64_t p = lit_start;
        while (p < end) {
            const int64_t run = end - p < MAX_LIT ? end - p : MAX_LIT;
            if (op + 1 + run > dst_cap) return false;
            dst[op++] = (uint8_t)(run - 1);
            memcpy(dst + op, src + p, (size_t)run);
            op += run;
            p += run;
        }
        return true;
    };
    while (ip + 2 < n) {
        const uint32_t v = (uint32_t)src[ip] << 16 | (uint32_t)src[ip + 1] << 8 | src[ip + 2];
        const uint32_t h = ((v * 2654435761u) >> (32 - HLOG)) & (HSIZE - 1);
        const int64_t ref = htab[h];
        htab[h] = ip;
        if (ref >= 0 && ip - ref <= MAX_OFF && src[ref] == src[ip] && src[ref + 1] == src[ip + 1] &&
            src[ref + 2] == src[ip + 2]) {
            int64_t len = 3;
            const int64_t max_len = n - ip < MAX_REF ? n - ip : MAX_REF;
            while (len < max_len && src[ref + len] == src[ip + len]) ++len;
            if (!flush_literals(ip)) return 0;
            const int64_t off = ip - ref - 1, l = len - 2;
            if (op + 3 > dst_cap) return 0;
            if (l < 7) dst[op++] = (uint8_t)((l << 5) | (off >> 8));
            else { dst[op++] = (uint8_t)((7 << 5) | (off >> 8)); dst[op++] = (uint8_t)(l - 7); }
            dst[op++] = (uint8_t)(off & 0xff);
            ip += len;
            lit_start = ip;
        } else {
            ++ip;
        }
    }
    if (!flush_literals(n)) return 0;
    return op;
}

// Chunks of the 'data' dataset of a --lowmem file straight from packed .jl rows (utmos/select.py:198-231): chunk i
// holds rows [i*chunk_rows, (i+1)*chunk_rows) as S bool bytes each, or S float32 GT*AF each when af != NULL; rows past
// n_rows are zero (h5py pads the last chunk).  Every chunk is unpacked and LZF-compressed by one of `threads` host
// threads into dst + i*chunk_nbytes: sizes_out[i] = compressed bytes and masks_out[i] = 0, or the raw bytes and
// masks_out[i] = 1 when LZF does not shrink it (what the hdf5 lzf filter does).
int utmos_h5_encode_chunks(const uint8_t *rows, int64_t n_rows, int64_t pitch, int64_t n_samples, const double *af,
                           int64_t chunk_rows, uint8_t *dst, int64_t *sizes_out, uint32_t *masks_out, int threads)
{
    if (!rows || n_rows < 0 || pitch < (n_samples + 7) / 8 || n_samples <= 0 || chunk_rows <= 0 || !dst || !sizes_out || !masks_out) {
        utmos::set_error("h5_encode_chunks: bad arguments");
        return UTMOS_E_ARG;
    }
    const int64_t n_chunks = (n_rows + chunk_rows - 1) / chunk_rows;
    const size_t item = af ? 4 : 1;
    const size_t chunk_nbytes = (size_t)chunk_rows * (size_t)n_samples * item;
    int hw = (int)std::thread::hardware_concurrency();
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(threads > 0 ? threads : (hw > 0 ? hw : 1), 64), n_chunks));
    std::atomic<int64_t> next(0);
    auto worker = [&]() {
        std::vector<uint8_t> dense(chunk_nbytes);
        for (;;) {
            const int64_t c = next.fetch_add(1);
            if (c >= n_chunks) return;
            memset(dense.data(), 0, chunk_nbytes);
            for (int64_t i = 0; i < chunk_rows; ++i) {
                const int64_t r = c * chunk_rows + i;
                if (r >= n_rows) break;
                const uint8_t *row = rows + (size_t)r * (size_t)pitch;
                if (af) {
                    float *out = reinterpret_cast<float *>(dense.data()) + (size_t)i * (size_t)n_samples;
                    const float v = (float)af[r];                     // (bool * float64).astype(float32), select.py:222
                    for (int64_t s = 0; s < n_samples; ++s)
                        if (row[s >> 3] & (0x80u >> (s & 7))) out[s] = v;
                } else {
                    uint8_t *out = dense.data() + (size_t)i * (size_t)n_samples;
                    for (int64_t s = 0; s < n_samples; ++s) out[s] = (row[s >> 3] >> (7 - (s & 7))) & 1u;
                }
            }
            uint8_t *slot = dst + (size_t)c * chunk_nbytes;
            const int64_t got = utmos_lzf_compress(dense.data(), (int64_t)chunk_nbytes, slot, (int64_t)chunk_nbytes - 1);
            if (got > 0) {
                sizes_out[c] = got;
                masks_out[c] = 0;
            } else {
                memcpy(slot, dense.data(), chunk_nbytes);
                sizes_out[c] = (int64_t)chunk_nbytes;
                masks_out[c] = 1;
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
    return UTMOS_OK;
}

}  // extern "C"
