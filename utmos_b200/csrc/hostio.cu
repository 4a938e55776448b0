// hostio.cu -- host-side byte codecs used by the hdf5 chunk streamer (no device code).
//
// hdf5 filter 32000 ("lzf", the codec h5py applies for compression="lzf", utmos/select.py:208-238) stores each
// chunk as one liblzf block.  Format restated from the published liblzf stream description:
//   ctrl < 32           : literal run of ctrl+1 bytes
//   ctrl >= 32          : back reference, len = ctrl>>5 (if 7: + next byte), then +2;
//                         offset = ((ctrl & 31) << 8 | next byte) + 1 bytes behind the write position
#include <stdint.h>
#include <string.h>

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

}  // extern "C"
