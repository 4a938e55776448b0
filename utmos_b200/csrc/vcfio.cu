// vcfio.cu -- host-side VCF text path of `utmos convert` (no device code): parallel BGZF / gzip inflate and a
// multi-threaded genotype tokenizer that fills the int8 GT[V][S][2] tensor the K1 kernel consumes.
//
// Replaces, on the host, what the reference gets from scikit-allel's Cython parser:
//   allel.read_vcf(in_file, fields=["calldata/GT", "samples"])        utmos/convert.py:50-53
// Semantics (pinned by tests against the pure-Python restatement utmos_b200/vcf.py, which in turn reproduces the
// reference's chunk{0,1}.vcf.gz -> chunk{0,1}.jl fixtures bit for bit): ploidy 2, a missing allele ('.', or the
// absent second allele of a haploid call) is -1, phasing is ignored, ploidy > 2 is truncated, a FORMAT without GT
// gives an all-missing row.
#include <stdint.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {

inline int n_threads(int want)
{
    int hw = (int)std::thread::hardware_concurrency();
    if (hw <= 0) hw = 1;
    return std::max(1, std::min(want > 0 ? want : hw, 64));
}

template <typename F>
void parallel_for(long long n, int threads, F fn)
{
    const int nt = (int)std::min<long long>(n_threads(threads), std::max<long long>(n, 1));
    if (nt <= 1) { fn(0, n); return; }
    std::vector<std::thread> pool;
    const long long per = (n + nt - 1) / nt;
    for (int t = 0; t < nt; ++t) {
        const long long b = t * per, e = std::min(n, b + per);
        if (b >= e) break;
        pool.emplace_back([=]() { fn(b, e); });
    }
    for (auto &t : pool) t.join();
}

struct BgzfBlock {
    long long off, size, isize, out;
};

// BGZF = concatenated gzip members of <= 64 KiB with an extra subfield 'BC' holding the member size - 1
bool bgzf_index(const uint8_t *src, long long len, std::vector<BgzfBlock> &blocks)
{
    long long p = 0, out = 0;
    while (p < len) {
        if (p + 18 > len || src[p] != 0x1f || src[p + 1] != 0x8b || src[p + 2] != 8 || !(src[p + 3] & 4)) return false;
        const int xlen = src[p + 10] | (src[p + 11] << 8);
        long long q = p + 12, xend = p + 12 + xlen;
        long long bsize = -1;
        while (q + 4 <= xend && xend <= len) {
            const int slen = src[q + 2] | (src[q + 3] << 8);
            if (src[q] == 'B' && src[q + 1] == 'C' && slen == 2 && q + 6 <= xend) bsize = (src[q + 4] | (src[q + 5] << 8)) + 1;
            q += 4 + slen;
        }
        if (bsize < 0 || p + bsize > len || bsize < 12 + xlen + 8) return false;
        const uint8_t *tail = src + p + bsize - 4;
        const long long isize = (long long)tail[0] | ((long long)tail[1] << 8) | ((long long)tail[2] << 16) | ((long long)tail[3] << 24);
        blocks.push_back(BgzfBlock{p + 12 + xlen, bsize - 12 - xlen - 8, isize, out});
        out += isize;
        p += bsize;
    }
    return !blocks.empty();
}

// window_bits -15: one raw deflate stream (a BGZF block); 31: a gzip file, possibly several members back to back
bool inflate_raw(const uint8_t *src, long long n, uint8_t *dst, long long cap, int window_bits, long long *got)
{
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, window_bits) != Z_OK) return false;
    long long ip = 0, op = 0;
    bool ok = false;
    for (;;) {
        if (zs.avail_in == 0 && ip < n) {                   // zlib counts in 32 bits: feed at most 1 GiB at a time
            const long long chunk = std::min<long long>(n - ip, 1ll << 30);
            zs.next_in = const_cast<Bytef *>(src + ip);
            zs.avail_in = (uInt)chunk;
            ip += chunk;
        }
        if (zs.avail_out == 0 && op < cap) {
            const long long chunk = std::min<long long>(cap - op, 1ll << 30);
            zs.next_out = dst + op;
            zs.avail_out = (uInt)chunk;
            op += chunk;
        }
        const int rc = inflate(&zs, Z_NO_FLUSH);
        if (rc == Z_STREAM_END) {
            if (window_bits > 15 && (zs.avail_in > 0 || ip < n)) {
                if (inflateReset(&zs) != Z_OK) break;       // next gzip member
                continue;
            }
            ok = true;
            break;
        }
        if (rc != Z_OK) break;                              // corrupt stream, or no progress possible
        if (zs.avail_in == 0 && ip >= n) break;             // truncated input
        if (zs.avail_out == 0 && op >= cap) break;          // output buffer full
    }
    *got = op - (long long)zs.avail_out;
    inflateEnd(&zs);
    return ok;
}

// the gi-th ':'-separated subfield of [b, e)
inline void subfield(const char *&b, const char *&e, int gi)
{
    const char *p = b;
    for (int k = 0; k < gi; ++k) {
        const char *c = (const char *)memchr(p, ':', (size_t)(e - p));
        if (!c) { b = e; return; }                            // fewer subfields than FORMAT promises: missing
        p = c + 1;
    }
    const char *c = (const char *)memchr(p, ':', (size_t)(e - p));
    b = p;
    if (c) e = c;
}

// int(text) if text.isdigit() else -1   (ASCII digits; values above 127 do not fit the int8 tensor -> -2)
inline int allele(const char *b, const char *e)
{
    if (b >= e) return -1;
    int v = 0;
    for (const char *p = b; p < e; ++p) {
        if (*p < '0' || *p > '9') return -1;
        v = v * 10 + (*p - '0');
        if (v > 127) return -2;
    }
    return v;
}

inline const char *find_sep(const char *b, const char *e)
{
    const char *c = (const char *)memchr(b, '|', (size_t)(e - b));
    if (!c) c = (const char *)memchr(b, '/', (size_t)(e - b));
    return c;
}

// utmos_b200/vcf.py:_parse_gt
inline bool parse_gt(const char *b, const char *e, int8_t *out)
{
    // fast path: d|d or d/d with single digits
    if (e - b == 3 && (b[1] == '|' || b[1] == '/') && b[0] >= '0' && b[0] <= '9' && b[2] >= '0' && b[2] <= '9') {
        out[0] = (int8_t)(b[0] - '0');
        out[1] = (int8_t)(b[2] - '0');
        return true;
    }
    const char *sep = find_sep(b, e);
    int a0, a1;
    if (!sep) {
        a0 = allele(b, e);
        a1 = -1;
    } else {
        const char *sb = sep + 1, *se = e;
        const char *nxt = find_sep(sb, se);
        if (nxt) se = nxt;                                     // ploidy > 2 is truncated
        a0 = allele(b, sep);
        a1 = allele(sb, se);
    }
    if (a0 == -2 || a1 == -2) return false;
    out[0] = (int8_t)a0;
    out[1] = (int8_t)a1;
    return true;
}

// one data line [b, e) (without the newline) -> row[S][2]; returns 0 ok, 1 too many sample columns, 2 allele > 127
int parse_line(const char *b, const char *e, long long S, int8_t *row)
{
    memset(row, 0xff, (size_t)S * 2);
    const char *p = b, *fmt_b = nullptr, *fmt_e = nullptr;
    for (int f = 0; f < 9; ++f) {
        const char *t = (const char *)memchr(p, '\t', (size_t)(e - p));
        if (f == 8) {
            fmt_b = p;
            fmt_e = t ? t : e;
            p = t ? t + 1 : nullptr;
            break;
        }
        if (!t) return 0;                                      // fewer than 9 columns: nothing to call
        p = t + 1;
    }
    if (!p) return 0;                                          // FORMAT is the last column: no sample columns
    // index of GT in FORMAT
    int gi = -1, k = 0;
    for (const char *q = fmt_b;; ++k) {
        const char *c = (const char *)memchr(q, ':', (size_t)(fmt_e - q));
        const char *qe = c ? c : fmt_e;
        if (qe - q == 2 && q[0] == 'G' && q[1] == 'T') { gi = k; break; }
        if (!c) break;
        q = c + 1;
    }
    if (gi < 0) return 0;                                      // no GT in FORMAT: all missing
    long long s = 0;
    for (;;) {
        // fast path, the bulk of every population VCF: "d|d<TAB>" / "d/d<TAB>" with GT first and nothing after it
        while (gi == 0 && e - p >= 4 && p[3] == '\t' && (p[1] == '|' || p[1] == '/') && (unsigned)(p[0] - '0') < 10u &&
               (unsigned)(p[2] - '0') < 10u && s < S) {
            row[s * 2] = (int8_t)(p[0] - '0');
            row[s * 2 + 1] = (int8_t)(p[2] - '0');
            ++s;
            p += 4;
        }
        const char *t = (const char *)memchr(p, '\t', (size_t)(e - p));
        const char *cb = p, *ce = t ? t : e;
        if (s >= S) return 1;
        if (gi > 0 || memchr(cb, ':', (size_t)(ce - cb))) subfield(cb, ce, gi);
        if (!parse_gt(cb, ce, row + s * 2)) return 2;
        ++s;
        if (!t) break;
        p = t + 1;
    }
    return 0;
}

}  // namespace

extern "C" {

// Uncompressed size of a gzip / BGZF buffer: exact for BGZF (sum of the blocks' ISIZE), and for a single-member
// gzip below 4 GiB (ISIZE of the trailer); -1 when the buffer is not gzip.  *is_bgzf_out tells which.
int64_t utmos_gz_size(const uint8_t *src, int64_t len, int *is_bgzf_out)
{
    if (is_bgzf_out) *is_bgzf_out = 0;
    if (!src || len < 18 || src[0] != 0x1f || src[1] != 0x8b) return -1;
    std::vector<BgzfBlock> blocks;
    if (bgzf_index(src, len, blocks)) {
        if (is_bgzf_out) *is_bgzf_out = 1;
        return blocks.back().out + blocks.back().isize;
    }
    const uint8_t *t = src + len - 4;
    return (int64_t)t[0] | ((int64_t)t[1] << 8) | ((int64_t)t[2] << 16) | ((int64_t)t[3] << 24);
}

// Inflate a whole gzip / BGZF buffer into dst.  BGZF blocks are independent deflate streams: `threads` host threads
// (0 = all cores) inflate them in parallel; plain gzip is inflated by one thread.  *dst_len_out = bytes written.
int utmos_gz_inflate(const uint8_t *src, int64_t len, uint8_t *dst, int64_t dst_cap, int64_t *dst_len_out, int threads)
{
    if (!src || !dst || !dst_len_out || len < 18) { utmos::set_error("gz_inflate: bad arguments"); return UTMOS_E_ARG; }
    std::vector<BgzfBlock> blocks;
    if (bgzf_index(src, len, blocks)) {
        const long long total = blocks.back().out + blocks.back().isize;
        if (total > dst_cap) { utmos::set_error("gz_inflate: output buffer too small"); return UTMOS_E_ARG; }
        std::atomic<int> bad(0);
        parallel_for((long long)blocks.size(), threads, [&](long long b, long long e) {
            for (long long i = b; i < e && !bad.load(); ++i) {
                const BgzfBlock &k = blocks[(size_t)i];
                long long got = 0;
                if (k.isize == 0) continue;
                if (!inflate_raw(src + k.off, k.size, dst + k.out, k.isize, -15, &got) || got != k.isize) bad.store(1);
            }
        });
        if (bad.load()) { utmos::set_error("gz_inflate: corrupt BGZF block"); return UTMOS_E_DATA; }
        *dst_len_out = total;
        return UTMOS_OK;
    }
    long long got = 0;
    if (!inflate_raw(src, len, dst, dst_cap, 31, &got)) { utmos::set_error("gz_inflate: corrupt gzip stream or output buffer too small"); return UTMOS_E_DATA; }
    *dst_len_out = got;
    return UTMOS_OK;
}

// Tokenise up to max_variants DATA lines of VCF text starting at text[0] (header lines, i.e. lines starting with
// '#', and empty lines are skipped) into gt_out[max_variants][n_samples][2] (int8, missing = -1).  Lines are found
// by one pass of memchr and parsed by `threads` host threads.  *n_out = variants written, *consumed_out = bytes of
// text consumed (always ends on a line boundary; a last line without '\n' is consumed only when `final` is set).
int utmos_vcf_parse_gt(const char *text, int64_t len, int64_t n_samples, int8_t *gt_out, int64_t max_variants,
                       int64_t *n_out, int64_t *consumed_out, int final, int threads)
{
    if (!text || len < 0 || n_samples <= 0 || !gt_out || max_variants < 0 || !n_out || !consumed_out) {
        utmos::set_error("vcf_parse_gt: bad arguments");
        return UTMOS_E_ARG;
    }
    std::vector<std::pair<const char *, const char *>> lines;
    lines.reserve((size_t)std::min<int64_t>(max_variants, 1 << 20));
    const char *p = text, *end = text + len;
    while (p < end && (int64_t)lines.size() < max_variants) {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        if (!nl && !final) break;
        const char *le = nl ? nl : end;
        if (le > p && le[-1] == '\r') --le;                   // CRLF files (Python's text mode drops the '\r' too)
        if (le > p && *p != '#') lines.emplace_back(p, le);
        p = nl ? nl + 1 : end;
    }
    // swallow header / empty lines that follow, so the caller's next call starts on a data line
    while (p < end && (*p == '\n')) ++p;
    std::atomic<int> bad(0);
    const long long S = n_samples;
    parallel_for((long long)lines.size(), threads, [&](long long b, long long e) {
        for (long long i = b; i < e; ++i) {
            const int rc = parse_line(lines[(size_t)i].first, lines[(size_t)i].second, S, gt_out + (size_t)i * (size_t)S * 2);
            if (rc) bad.store(rc);
        }
    });
    if (bad.load() == 1) { utmos::set_error("vcf_parse_gt: a line has more sample columns than the #CHROM header"); return UTMOS_E_DATA; }
    if (bad.load() == 2) { utmos::set_error("vcf_parse_gt: allele index above 127"); return UTMOS_E_DATA; }
    *n_out = (int64_t)lines.size();
    *consumed_out = (int64_t)(p - text);
    return UTMOS_OK;
}

}  // extern "C"
