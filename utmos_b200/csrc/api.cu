// api.cu -- context management and the extern "C" surface declared in include/utmos_b200.h.
//
// HBM layout owned by a context (S samples, V informative rows):
//   rows   uint32 [V][pitchW]        variant-major bit matrix, pitchW = ceil(S/32) rounded up to 4 words
//   cols   uint32 [S32][colPitchW]   sample-major copy (built at finalize when it fits), colPitchW =
//                                    ceil(V/32) rounded up to 8 words
//   af     double [V]                per-row allele frequency (AF flavours only)
//   q_lo/q_hi uint64 [V]             AF * 2^scale as two limbs of L bits (AF flavours only)
//   live0/live uint32 [colPitchW]    scoring rows at step 0 / rows not yet covered
//   gain0_* / gain_*  [S]            per-sample gains at step 0 / current
//   mask u8[S], weights f64[S], out_{idx,new,score}[S], SelState
#include <fcntl.h>
#include <math.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>

#include "common.cuh"

namespace utmos {

static thread_local std::string g_error;

void set_error(const std::string &msg) { g_error = msg; }

int cuda_fail(cudaError_t err, const char *what, const char *file, int line)
{
    char buf[512];
    snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d: %s", (int)err, cudaGetErrorString(err), file, line, what);
    g_error = buf;
    cudaGetLastError();   // clear the sticky-less error state
    if (err == cudaErrorMemoryAllocation) return UTMOS_E_NOMEM;
    if (err == cudaErrorNoDevice || err == cudaErrorInsufficientDriver) return UTMOS_E_NOGPU;
    return UTMOS_E_CUDA;
}

namespace {
// UTMOS_B200_TRACE=1: host-side wall time of the API sub-steps on stderr
struct Trace {
    bool on;
    std::chrono::steady_clock::time_point t0;
    const char *what;
    explicit Trace(const char *w) : on(getenv("UTMOS_B200_TRACE") != nullptr), t0(std::chrono::steady_clock::now()), what(w) {}
    void lap(const char *label)
    {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[utmos_b200 trace] %s: %s %.3f ms\n", what, label,
                std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};
constexpr size_t kStageBytes = 64ull << 20;      // per staging buffer (two pinned host + two device)
constexpr int kGraphSteps = 32;                  // step pairs per CUDA graph replay
// edge-list entries (16 B each; 32 B for AF flavours).  2^24 until the entry-divided cluster tail existed; with it a heavy
// tail step costs ~5 us where a head step costs 20-60 us, so the hand-over may come as early as the lists fit 1-2 GB
// (8 x 1,103,547 rows on one GPU: 36.5 -> 32.1 ms; the 1kGP shape hands over on the rows-per-pick rule either way)
constexpr unsigned long long kListBudget = 1ull << 26;
constexpr unsigned long long kListBudgetWide = 1ull << 28;   // S > 65,535: the rare variants of a 100k-sample cohort alone
                                                             // hold more than 2^24 bits, and the head costs ~20-100 us per step
enum { T_H2D = 0, T_INGEST = 1, T_TRANSPOSE = 2, T_GAIN = 3, T_SELECT = 4, T_COUNT = 5 };
// parts of the select loop (host clock between stream synchronisations): head launches, hand-over to the lists, tail
enum { P_HEAD = 0, P_HANDOVER = 1, P_TAIL = 2, P_COUNT = 3 };

struct Pending {
    int cat;
    cudaEvent_t a, b;
};
}  // namespace

}  // namespace utmos

using namespace utmos;

struct utmos_ctx {
    int device = 0;
    int n_sms = 0;
    long long S = 0;
    int nW = 0, pitchW = 0;
    int af_mode = UTMOS_AF_NONE;
    uint32_t flags = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;

    uint32_t *d_rows = nullptr;
    double *d_af = nullptr;
    long long rows_cap = 0;
    long long rows_upper = 0;          // upper bound of rows stored (kept rows are only known on the device)
    long long *d_nrows = nullptr;      // [0] rows stored, [1] scratch
    IngestScratch scratch;

    void *h_stage[2] = {nullptr, nullptr};
    void *d_stage[2] = {nullptr, nullptr};
    double *h_af_stage[2] = {nullptr, nullptr};
    double *d_af_stage[2] = {nullptr, nullptr};
    long long af_stage_rows = 0;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr};
    cudaEvent_t ev_consumed[2] = {nullptr, nullptr};
    bool stage_used[2] = {false, false};
    int stage_next = 0;
    bool owns_pinned_cache = false;

    bool finalized = false;
    long long V = 0, colPitchW = 0;
    int L = 0, scale = 0;
    uint32_t *d_cols = nullptr, *d_live = nullptr, *d_live0 = nullptr;
    unsigned int *d_gain_cnt = nullptr, *d_gain0_cnt = nullptr, *d_var_count = nullptr;
    unsigned long long *d_gain_lo = nullptr, *d_gain_hi = nullptr, *d_gain0_lo = nullptr, *d_gain0_hi = nullptr;
    unsigned long long *d_q_lo = nullptr, *d_q_hi = nullptr;
    uint8_t *d_mask = nullptr;
    uint32_t *d_selw = nullptr;        // bitmask of the samples that were selectable at select_begin
    double *d_weights = nullptr;
    bool has_weights = false;
    long long *d_out_idx = nullptr, *d_out_new = nullptr;
    double *d_out_score = nullptr, *d_dbg_score = nullptr;
    long long *d_out_time = nullptr;
    long long *d_dbg = nullptr;
    uint4 *d_lists[2] = {nullptr, nullptr};            // edge lists, double buffered for re-compaction
    unsigned int *d_list_off[2] = {nullptr, nullptr}, *d_list_len[2] = {nullptr, nullptr}, *d_cursor = nullptr, *d_pool_cursor = nullptr;
    unsigned short *d_pool = nullptr;
    size_t pool_cap = 0;
    size_t lists_cap[2] = {0, 0};      // uint4 units allocated
    int lists_cur = 0;
    unsigned long long lists_total = 0;   // entries in the current lists (live bits when they were built)
    bool lists_valid = false;
    int dbg_time = 0;
    // multi-GPU (rows sharded over ranks)
    int mg_rank = 0, mg_world = 1;
    long long global_rows = -1;        // informative rows summed over the ranks (set before finalize)
    void *mg_block = nullptr;          // IPC-exported exchange block: inboxes + flags
    size_t mg_bytes = 0;
    void *mg_peer[kMaxRanks] = {nullptr};
    unsigned int *d_delta_cnt = nullptr;
    unsigned long long *d_delta_lo = nullptr, *d_delta_hi = nullptr;
    int mg_grid = 0, mg_block_threads = 0;
    // hand-over to the replicated tail (mgpu.cu): merged numbering of the rows of all ranks
    long long mg_row_base = 0;         // first merged row id of this rank (multiple of 32)
    long long mg_merged_rows = 0;      // sum over ranks of their rows rounded up to 32
    bool mg_allow_tail = false;        // every rank has a sample-major copy (set by utmos_mgpu_layout)
    bool mg_tail = false;              // the merged lists are built: every rank runs the same tail kernel
    unsigned long long mg_list_cap = 0;   // entries of the merged lists region of the exchange block
    unsigned int *d_lcnt = nullptr, *d_my_base = nullptr, *d_pool_base = nullptr;
    unsigned int *d_local0_cnt = nullptr, *d_local_cnt = nullptr;          // this rank's share of the gains (step 0 / now)
    unsigned long long *d_local0_lo = nullptr, *d_local0_hi = nullptr, *d_local_lo = nullptr, *d_local_hi = nullptr;
    bool lists_external = false;       // d_lists[0] / d_pool live inside the exchange block (not owned)
    unsigned int tail_rows = 2048;        // hand over to the list-driven tail once picks cover fewer rows than this
    unsigned int tie_row_cap = 0;         // UTMOS_OPT_TIE_ROW_CAP
    bool ref_hybrid = false;              // UTMOS_F_REF_TIES without UTMOS_F_STEP_KERNELS: step kernels, then the REFT tail
    unsigned long long list_budget = 0;   // edge-list entries the tail may be built from (0 = kListBudget / kListBudgetWide)
    unsigned int tail_heavy_rows = 0xffffffffu;   // (0xffffffff = default: 768 in count mode, 1 = every tail step with AF or a state that does not fit one SM)
                                          // list-driven tail: picks that cover at least this many rows are run by the entry-divided
                                          // 16-CTA cluster kernel (state sliced over its shared memories), lighter ones on one SM (0 = never)
    unsigned int tail_single_rows = 0;    // > 0: 8-CTA owner-computes flavour of the tail until picks cover fewer rows than this
                                          // (measured on the 1kGP shape: not faster than one CTA, so off by default)
    uint32_t *d_newmask = nullptr;        // rows newly covered by the pick that ended a head launch (cover_decrement_kernel)
    uint32_t *d_live_priv = nullptr;      // private live masks of the cluster flavour when they do not fit in shared memory
    size_t live_priv_bytes = 0;
    unsigned long long tail_budget = 0;   // handed to the head kernels while the tail flavour waits for sparsity
    unsigned long long total_bits = 0;    // set bits of the scoring rows at step 0
    long long regain_rows = -1;        // -1 = default heuristic
    SelState *d_state = nullptr;
    unsigned int *d_bar = nullptr;
    ArgPartial *d_partials = nullptr;
    int grid = 0, block = 0;
    bool selecting = false;
    cudaGraphExec_t graph_exec = nullptr;
    int flavour_used = 0;
    int cluster = 0;
    unsigned int af_inexact = 0;

    int n_launch = 0;
    double ms[T_COUNT] = {0, 0, 0, 0, 0};
    double part_ms[P_COUNT] = {0.0, 0.0, 0.0};
    cudaEvent_t ev_handover[2] = {nullptr, nullptr};
    bool handover_unread = false;
    std::vector<Pending> pending;
    size_t dev_bytes = 0;
};

namespace {

// Device memory comes from the device's stream-ordered pool (cudaMallocAsync) with the release threshold
// lifted, so a process that runs many selections re-uses its buffers instead of paying cudaMalloc/cudaFree
// for hundreds of MB each time.  Allocations are ordered on c->stream; callers that hand a fresh buffer to
// the copy stream synchronise c->stream first (ensure_stage).
int dev_alloc(utmos_ctx *c, void **p, size_t bytes)
{
    if (bytes == 0) bytes = 16;
    UT_CUDA(cudaMallocAsync(p, bytes, c->stream));
    c->dev_bytes += bytes;
    return UTMOS_OK;
}

template <typename T>
void dev_free(utmos_ctx *c, T *&p, size_t bytes)
{
    if (p) {
        cudaFreeAsync(p, c->stream);
        c->dev_bytes -= std::min(c->dev_bytes, bytes ? bytes : (size_t)16);
        p = nullptr;
    }
}

// Big blocks (bit matrices, edge lists, staging) come from cudaMalloc and are parked in a small process-wide
// cache when a context lets go of them: cudaMalloc / cudaFree of hundreds of MB cost milliseconds and cudaFree
// synchronises the device, which a service that runs selection after selection must not pay every time.
struct BigBlock {
    void *ptr;
    size_t bytes;
    int device;
};
std::vector<BigBlock> g_big_cache;
constexpr size_t kBigMin = 8ull << 20;
constexpr size_t kBigCacheMax = 24;

int big_alloc(utmos_ctx *c, void **p, size_t bytes);
void big_free_raw(int device, void *ptr, size_t bytes);
template <typename T>
void big_free(utmos_ctx *c, T *&p, size_t bytes);

// pinned staging buffers are expensive to create (page locking): keep two per process and lend them out
struct PinnedCache {
    void *buf[2] = {nullptr, nullptr};
    bool busy = false;
};
PinnedCache g_pinned;

int pinned_acquire(utmos_ctx *c);
void pinned_release(utmos_ctx *c);
void mg_block_release(void *ptr);

void t_begin(utmos_ctx *c, int cat, cudaStream_t s)
{
    Pending p;
    p.cat = cat;
    cudaEventCreate(&p.a);
    cudaEventCreate(&p.b);
    cudaEventRecord(p.a, s);
    c->pending.push_back(p);
}

void t_end(utmos_ctx *c, cudaStream_t s) { cudaEventRecord(c->pending.back().b, s); }

void t_resolve(utmos_ctx *c)      // call only after the streams have been synchronised
{
    for (auto &p : c->pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) c->ms[p.cat] += ms;
        cudaEventDestroy(p.a);
        cudaEventDestroy(p.b);
    }
    c->pending.clear();
    cudaGetLastError();
}

int sync_all(utmos_ctx *c)
{
    UT_CUDA(cudaStreamSynchronize(c->copy_stream));
    UT_CUDA(cudaStreamSynchronize(c->stream));
    t_resolve(c);
    return UTMOS_OK;
}

int grow_rows(utmos_ctx *c, long long need)
{
    if (need <= c->rows_cap) return UTMOS_OK;
    long long cap = std::max(need, c->rows_cap + c->rows_cap / 2);
    cap = std::max(cap, 1024ll);
    uint32_t *nr = nullptr;
    double *na = nullptr;
    const size_t row_bytes = (size_t)c->pitchW * 4;
    UT_TRY(big_alloc(c, (void **)&nr, (size_t)cap * row_bytes));
    if (c->af_mode != UTMOS_AF_NONE) UT_TRY(dev_alloc(c, (void **)&na, (size_t)cap * 8));
    if (c->d_rows) {
        // rows already ingested are copied over stream-ordered behind the ingest kernels
        UT_CUDA(cudaMemcpyAsync(nr, c->d_rows, (size_t)c->rows_upper * row_bytes, cudaMemcpyDeviceToDevice, c->stream));
        if (na) UT_CUDA(cudaMemcpyAsync(na, c->d_af, (size_t)c->rows_upper * 8, cudaMemcpyDeviceToDevice, c->stream));
        UT_CUDA(cudaStreamSynchronize(c->stream));
        big_free(c, c->d_rows, (size_t)c->rows_cap * row_bytes);
        dev_free(c, c->d_af, (size_t)c->rows_cap * 8);
    }
    c->d_rows = nr;
    c->d_af = na;
    c->rows_cap = cap;
    return UTMOS_OK;
}

int big_alloc(utmos_ctx *c, void **p, size_t bytes)
{
    if (bytes < kBigMin) return dev_alloc(c, p, bytes);
    int best = -1;
    for (size_t i = 0; i < g_big_cache.size(); ++i) {
        const BigBlock &b = g_big_cache[i];
        if (b.device == c->device && b.bytes >= bytes && b.bytes <= bytes + bytes / 4 &&
            (best < 0 || b.bytes < g_big_cache[best].bytes))
            best = (int)i;
    }
    if (best >= 0) {
        *p = g_big_cache[best].ptr;
        g_big_cache.erase(g_big_cache.begin() + best);
    } else {
        cudaError_t e = cudaMalloc(p, bytes);
        if (e == cudaErrorMemoryAllocation && !g_big_cache.empty()) {
            cudaGetLastError();
            for (auto &b : g_big_cache) cudaFree(b.ptr);      // give parked blocks back and retry
            g_big_cache.clear();
            e = cudaMalloc(p, bytes);
        }
        UT_CUDA(e);
    }
    c->dev_bytes += bytes;
    return UTMOS_OK;
}

void big_free_raw(int device, void *ptr, size_t bytes)
{
    if (!ptr) return;
    if (g_big_cache.size() >= kBigCacheMax) {
        cudaFree(g_big_cache.front().ptr);
        g_big_cache.erase(g_big_cache.begin());
    }
    g_big_cache.push_back(BigBlock{ptr, bytes, device});
}

// the caller guarantees that no work touching the block is still in flight (streams synchronised)
template <typename T>
void big_free(utmos_ctx *c, T *&p, size_t bytes)
{
    if (!p) return;
    if (bytes < kBigMin) { dev_free(c, p, bytes); return; }
    c->dev_bytes -= std::min(c->dev_bytes, bytes);
    // a block handed out for `bytes` may be larger: the cache entry must remember the real size, which the
    // cache itself cannot know here -> store the requested size (it only ever satisfies requests <= that)
    big_free_raw(c->device, (void *)p, bytes);
    p = nullptr;
}

int pinned_acquire(utmos_ctx *c)
{
    if (c->h_stage[0]) return UTMOS_OK;
    if (!g_pinned.busy) {
        for (int i = 0; i < 2; ++i)
            if (!g_pinned.buf[i]) UT_CUDA(cudaMallocHost(&g_pinned.buf[i], kStageBytes));
        g_pinned.busy = true;
        c->owns_pinned_cache = true;
        c->h_stage[0] = g_pinned.buf[0];
        c->h_stage[1] = g_pinned.buf[1];
    } else {
        for (int i = 0; i < 2; ++i) UT_CUDA(cudaMallocHost(&c->h_stage[i], kStageBytes));
    }
    return UTMOS_OK;
}

void pinned_release(utmos_ctx *c)
{
    if (c->owns_pinned_cache) {
        g_pinned.busy = false;
        c->owns_pinned_cache = false;
    } else {
        for (int i = 0; i < 2; ++i)
            if (c->h_stage[i]) cudaFreeHost(c->h_stage[i]);
    }
    c->h_stage[0] = c->h_stage[1] = nullptr;
}

int ensure_stage(utmos_ctx *c, long long af_rows, bool need_host)
{
    bool fresh = false;
    for (int i = 0; i < 2; ++i) {
        if (!c->d_stage[i]) {
            UT_TRY(big_alloc(c, &c->d_stage[i], kStageBytes));
            UT_CUDA(cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
            UT_CUDA(cudaEventCreateWithFlags(&c->ev_consumed[i], cudaEventDisableTiming));
            fresh = true;
        }
    }
    if (need_host) UT_TRY(pinned_acquire(c));
    if (af_rows > c->af_stage_rows) {
        UT_CUDA(cudaStreamSynchronize(c->stream));
        UT_CUDA(cudaStreamSynchronize(c->copy_stream));
        for (int i = 0; i < 2; ++i) {
            dev_free(c, c->d_af_stage[i], (size_t)c->af_stage_rows * 8);
            if (c->h_af_stage[i]) cudaFreeHost(c->h_af_stage[i]);
            UT_TRY(dev_alloc(c, (void **)&c->d_af_stage[i], (size_t)af_rows * 8));
            UT_CUDA(cudaMallocHost((void **)&c->h_af_stage[i], (size_t)af_rows * 8));
        }
        c->af_stage_rows = af_rows;
        fresh = true;
    }
    if (fresh) UT_CUDA(cudaStreamSynchronize(c->stream));   // pool allocations are ordered on c->stream
    return UTMOS_OK;
}

bool is_pinned(const void *p)
{
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return attr.type == cudaMemoryTypeHost;
}

// pageable -> pinned staging with several host threads: one core copies at 10-14 GB/s, which is a quarter of what the
// PCIe link takes (joblib.load hands the CLI pageable memory, so this is the path `utmos select a.jl` really uses)
void staged_memcpy(void *dst, const void *src, size_t bytes)
{
    static const int n_threads = std::max(1, std::min(8, (int)std::thread::hardware_concurrency() / 2));
    if (bytes < (8u << 20) || n_threads == 1) { memcpy(dst, src, bytes); return; }
    const size_t piece = (bytes / (size_t)n_threads + 4095) & ~(size_t)4095;
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) {
        const size_t off = piece * (size_t)t;
        if (off >= bytes) break;
        const size_t n = std::min(piece, bytes - off);
        pool.emplace_back([=]() { memcpy((char *)dst + off, (const char *)src + off, n); });
    }
    memcpy(dst, src, std::min(piece, bytes));
    for (auto &th : pool) th.join();
}

// host chunk pipeline: (pageable -> pinned staging ->) cudaMemcpyAsync on the copy stream -> ingest kernels
int append_host(utmos_ctx *c, int kind, const void *src, long long n_rows, long long pitch_in, const double *af)
{
    if (c->finalized) { set_error("append after finalize"); return UTMOS_E_ARG; }
    if (n_rows < 0 || (n_rows > 0 && !src)) { set_error("append: bad rows pointer / count"); return UTMOS_E_ARG; }
    if (n_rows == 0) return UTMOS_OK;
    const bool want_af = c->af_mode != UTMOS_AF_NONE;
    if (kind == RAW_PACKED_MSB && want_af && !af) { set_error("append_packed: AF required for an AF context"); return UTMOS_E_ARG; }
    UT_TRY(grow_rows(c, c->rows_upper + n_rows));
    const long long chunk_rows = std::max(1ll, (long long)(kStageBytes / (size_t)pitch_in));
    if ((size_t)pitch_in > kStageBytes) { set_error("append: one row exceeds the staging buffer"); return UTMOS_E_ARG; }
    const bool need_af_stage = (kind == RAW_PACKED_MSB && want_af) || kind == RAW_DENSE_F32;
    const bool src_pinned = is_pinned(src);
    const bool af_pinned = af && is_pinned(af);
    UT_TRY(ensure_stage(c, need_af_stage ? std::min(chunk_rows, n_rows) : 0, !src_pinned));
    for (long long r0 = 0; r0 < n_rows; r0 += chunk_rows) {
        const long long n = std::min(chunk_rows, n_rows - r0);
        const int b = c->stage_next;
        c->stage_next ^= 1;
        const size_t bytes = (size_t)n * (size_t)pitch_in;
        const uint8_t *chunk = (const uint8_t *)src + (size_t)r0 * (size_t)pitch_in;
        if (c->stage_used[b]) {
            UT_CUDA(cudaEventSynchronize(c->ev_consumed[b]));                 // pinned host buffer reusable
            UT_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_consumed[b], 0));
        }
        t_begin(c, T_H2D, c->copy_stream);
        if (src_pinned) {
            UT_CUDA(cudaMemcpyAsync(c->d_stage[b], chunk, bytes, cudaMemcpyHostToDevice, c->copy_stream));
        } else {
            staged_memcpy(c->h_stage[b], chunk, bytes);
            UT_CUDA(cudaMemcpyAsync(c->d_stage[b], c->h_stage[b], bytes, cudaMemcpyHostToDevice, c->copy_stream));
        }
        const double *d_af_chunk = nullptr;
        if (kind == RAW_PACKED_MSB && want_af) {
            if (af_pinned) {
                UT_CUDA(cudaMemcpyAsync(c->d_af_stage[b], af + r0, (size_t)n * 8, cudaMemcpyHostToDevice, c->copy_stream));
            } else {
                memcpy(c->h_af_stage[b], af + r0, (size_t)n * 8);
                UT_CUDA(cudaMemcpyAsync(c->d_af_stage[b], c->h_af_stage[b], (size_t)n * 8, cudaMemcpyHostToDevice,
                                        c->copy_stream));
            }
            d_af_chunk = c->d_af_stage[b];
        } else if (kind == RAW_DENSE_F32) {
            d_af_chunk = c->d_af_stage[b];      // filled by the flag kernel
        }
        t_end(c, c->copy_stream);
        UT_CUDA(cudaEventRecord(c->ev_copied[b], c->copy_stream));
        UT_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copied[b], 0));
        t_begin(c, T_INGEST, c->stream);
        UT_TRY(launch_ingest(c->stream, c->scratch, kind, c->d_stage[b], n, pitch_in, d_af_chunk, (int)c->S, c->pitchW,
                             c->d_rows, want_af ? c->d_af : nullptr, c->d_nrows, &c->n_launch));
        t_end(c, c->stream);
        UT_CUDA(cudaEventRecord(c->ev_consumed[b], c->stream));
        c->stage_used[b] = true;
    }
    c->rows_upper += n_rows;
    return UTMOS_OK;
}

void free_select_state(utmos_ctx *c)
{
    const size_t S = (size_t)c->S;
    if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
    big_free(c, c->d_cols, (size_t)((c->S + 31) / 32 * 32) * (size_t)c->colPitchW * 4);
    dev_free(c, c->d_live, (size_t)c->colPitchW * 4);
    dev_free(c, c->d_live0, (size_t)c->colPitchW * 4);
    dev_free(c, c->d_gain_cnt, S * 4);
    dev_free(c, c->d_gain0_cnt, S * 4);
    dev_free(c, c->d_var_count, S * 4);
    dev_free(c, c->d_gain_lo, S * 8);
    dev_free(c, c->d_gain_hi, S * 8);
    dev_free(c, c->d_gain0_lo, S * 8);
    dev_free(c, c->d_gain0_hi, S * 8);
    dev_free(c, c->d_q_lo, (size_t)c->V * 8);
    dev_free(c, c->d_q_hi, (size_t)c->V * 8);
    dev_free(c, c->d_mask, S);
    dev_free(c, c->d_selw, (size_t)c->nW * 4);
    dev_free(c, c->d_weights, S * 8);
    dev_free(c, c->d_out_idx, S * 8);
    dev_free(c, c->d_out_new, S * 8);
    dev_free(c, c->d_out_score, S * 8);
    dev_free(c, c->d_out_time, S * 8);
    dev_free(c, c->d_dbg, 128);
    if (c->lists_external) {             // regions of the exchange block, not allocations of their own
        c->d_lists[0] = nullptr;
        c->d_pool = nullptr;
        c->lists_external = false;
    }
    dev_free(c, c->d_live_priv, c->live_priv_bytes);
    dev_free(c, c->d_newmask, (size_t)c->colPitchW * 4);
    c->live_priv_bytes = 0;
    dev_free(c, c->d_local0_cnt, S * 4);
    dev_free(c, c->d_local_cnt, S * 4);
    dev_free(c, c->d_local0_lo, S * 8);
    dev_free(c, c->d_local0_hi, S * 8);
    dev_free(c, c->d_local_lo, S * 8);
    dev_free(c, c->d_local_hi, S * 8);
    dev_free(c, c->d_lcnt, S * 4);
    dev_free(c, c->d_my_base, S * 4);
    dev_free(c, c->d_pool_base, 16);
    for (int i = 0; i < 2; ++i) {
        big_free(c, c->d_lists[i], c->lists_cap[i] * 16);
        dev_free(c, c->d_list_off[i], S * 4);
        dev_free(c, c->d_list_len[i], S * 4);
        c->lists_cap[i] = 0;
    }
    dev_free(c, c->d_cursor, S * 4);
    dev_free(c, c->d_delta_cnt, S * 4);
    dev_free(c, c->d_delta_lo, S * 8);
    dev_free(c, c->d_delta_hi, S * 8);
    for (int i = 0; i < kMaxRanks; ++i) c->mg_peer[i] = nullptr;     // mappings stay in the process-wide cache
    if (c->mg_block) { mg_block_release(c->mg_block); c->mg_block = nullptr; }
    dev_free(c, c->d_pool_cursor, 16);
    big_free(c, c->d_pool, c->pool_cap * (c->S > 65535 ? 4 : 2));
    c->pool_cap = 0;
    c->lists_valid = false;
    dev_free(c, c->d_dbg_score, S * 8);
    dev_free(c, c->d_bar, 64);
    dev_free(c, c->d_partials, sizeof(ArgPartial) * 2048);
}

struct MgLayout {
    size_t off_cnt, off_lo, off_hi, off_flags, small_bytes;     // per-step exchange: inboxes + flags (zeroed at export)
    size_t off_live, off_lists, off_pool, bytes;                 // merged tail structures (replicated on every rank)
    size_t live_words, pool_cap;
};
MgLayout mg_layout(size_t S, int world, bool af, unsigned long long list_cap, long long merged_rows)
{
    const size_t pool_elem = S > 65535 ? 4 : 2;           // carrier ids: uint16, or uint32 for wide cohorts
    MgLayout l;
    size_t off = 0;
    auto take = [&](size_t b) { const size_t o = off; off = (off + b + 255) / 256 * 256; return o; };
    l.off_cnt = take(2 * (size_t)world * S * 4);
    l.off_lo = take(af ? 2 * (size_t)world * S * 8 : 0);
    l.off_hi = take(af ? 2 * (size_t)world * S * 8 : 0);
    l.off_flags = take((size_t)kMaxRanks * 8);
    l.small_bytes = off;
    l.live_words = list_cap ? (size_t)std::max(8ll, (merged_rows / 32 + 7) / 8 * 8) : 0;
    l.pool_cap = list_cap ? (size_t)(mgpu_pool_share(list_cap) + 64ull * (size_t)world) : 0;
    l.off_live = take(l.live_words * 4);
    l.off_lists = take((size_t)list_cap * (af ? 32 : 16));
    l.off_pool = take(l.pool_cap * pool_elem);
    l.bytes = off;
    return l;
}

// Exchange blocks are cudaMalloc'ed once and kept for the life of the process: creating and IPC-mapping hundreds
// of MB per selection costs milliseconds and cudaFree synchronises the device.  A block handed back by a context
// is reused by the next one that fits; peers keep their mapping of it (keyed by the 64-byte IPC handle).
struct MgBlock {
    int device;
    size_t bytes;
    void *ptr;
    cudaIpcMemHandle_t handle;
    bool busy;
};
std::vector<MgBlock> g_mg_blocks;
struct MgPeerMap {
    int device;
    cudaIpcMemHandle_t handle;
    void *ptr;
};
std::vector<MgPeerMap> g_mg_peers;

int mg_block_acquire(int device, size_t bytes, void **ptr_out, cudaIpcMemHandle_t *handle_out)
{
    int best = -1;
    for (size_t i = 0; i < g_mg_blocks.size(); ++i) {
        const MgBlock &b = g_mg_blocks[i];
        if (!b.busy && b.device == device && b.bytes >= bytes && (best < 0 || b.bytes < g_mg_blocks[best].bytes)) best = (int)i;
    }
    if (best < 0) {
        MgBlock b;
        b.device = device;
        b.bytes = bytes;
        b.busy = false;
        UT_CUDA(cudaMalloc(&b.ptr, bytes));                  // plain cudaMalloc: IPC handles need it
        UT_CUDA(cudaIpcGetMemHandle(&b.handle, b.ptr));
        g_mg_blocks.push_back(b);
        best = (int)g_mg_blocks.size() - 1;
    }
    g_mg_blocks[best].busy = true;
    *ptr_out = g_mg_blocks[best].ptr;
    *handle_out = g_mg_blocks[best].handle;
    return UTMOS_OK;
}

void mg_block_release(void *ptr)
{
    for (auto &b : g_mg_blocks)
        if (b.ptr == ptr) b.busy = false;
}

int mg_peer_map(int device, const cudaIpcMemHandle_t &h, void **ptr_out)
{
    for (auto &m : g_mg_peers)
        if (m.device == device && memcmp(&m.handle, &h, sizeof(h)) == 0) { *ptr_out = m.ptr; return UTMOS_OK; }
    MgPeerMap m;
    m.device = device;
    m.handle = h;
    UT_CUDA(cudaIpcOpenMemHandle(&m.ptr, h, cudaIpcMemLazyEnablePeerAccess));
    g_mg_peers.push_back(m);
    *ptr_out = m.ptr;
    return UTMOS_OK;
}


SelParams make_params(const utmos_ctx *c, bool step0)
{
    SelParams p;
    memset(&p, 0, sizeof(p));
    p.rows = c->d_rows;
    p.cols = c->V > 0 ? c->d_cols : nullptr;
    p.live = step0 ? c->d_live0 : c->d_live;
    p.gain_cnt = step0 ? c->d_gain0_cnt : c->d_gain_cnt;
    p.gain_lo = step0 ? c->d_gain0_lo : c->d_gain_lo;
    p.gain_hi = step0 ? c->d_gain0_hi : c->d_gain_hi;
    p.q_lo = c->d_q_lo;
    p.q_hi = c->d_q_hi;
    p.mask = c->d_mask;
    p.selw = c->d_selw;
    p.weights = c->has_weights ? c->d_weights : nullptr;
    p.af_vals = c->d_af;
    p.ref_ties = ((c->flags & UTMOS_F_REF_TIES) && c->mg_world <= 1) ? 1 : 0;
    p.tie_row_cap = c->tie_row_cap;
    p.af_f32 = c->af_mode == UTMOS_AF_F32 ? 1 : 0;
    p.out_idx = c->d_out_idx;
    p.out_new = c->d_out_new;
    p.out_score = c->d_out_score;
    p.out_time = c->d_out_time;
    p.dbg = c->d_dbg;
    p.lists = c->d_lists[c->lists_cur];
    p.list_off = c->d_list_off[c->lists_cur];
    p.list_len = c->d_list_len[c->lists_cur];
    p.pool = c->d_pool;
    {
        // a pick that newly covers this many rows is cheaper to absorb by one streaming recompute of all gains
        // (S columns, ~V*S/8 bytes) than by one atomic per set bit of those rows
        // (measured on the 1kGP shape: one recompute = 90 us; subtracting a pick of V/170 rows costs the same)
        long long thr = c->regain_rows >= 0 ? c->regain_rows : std::max(4096ll, c->V / 170);
        if (!c->d_cols || (c->flags & UTMOS_F_STEP_KERNELS)) thr = 0;
        p.regain_rows = (unsigned int)std::min(thr, 0xffffffffll);
    }
    p.tail_budget = c->tail_budget;
    // N ranks: N times the rows per pick at the same sparsity (measured on 4 x B200: 1536 per rank beats 2048)
    p.tail_rows = c->mg_world > 1 && c->tail_rows == 2048u ? 1536u * (unsigned int)c->mg_world : c->tail_rows;
    // one GPU, many rows: a pick covers rows in proportion to V, so the hand-over point moves with it (8 x 1,103,547 rows
    // x 2,504 samples: 46.5 ms with 16,384 against 73.1 ms with 2,048; the 1kGP shape keeps 2,048)
    if (c->mg_world == 1 && c->tail_rows == 2048u && c->S <= 65535)
        p.tail_rows = (unsigned int)std::min<long long>(std::max<long long>(2048, c->V / 540), 1 << 20);
    p.dbg_time = c->dbg_time;
    p.dsmem_gains = (c->flags & UTMOS_F_DSMEM_GAINS) ? 1 : 0;
    p.st = c->d_state;
    p.V = c->V;
    p.colPitchW = c->colPitchW;
    p.S = (int)c->S;
    p.pitchW = c->pitchW;
    p.nW = c->nW;
    p.L = c->L;
    p.scale = c->scale;
    p.af = c->af_mode != UTMOS_AF_NONE;
    if (c->mg_tail) {
        // replicated tail over the merged structures of all ranks (mgpu.cu): global row numbering, no bit matrix
        const MgLayout l = mg_layout((size_t)c->S, c->mg_world, p.af != 0, c->mg_list_cap, c->mg_merged_rows);
        p.rows = nullptr;
        p.cols = nullptr;
        p.q_lo = p.q_hi = nullptr;
        p.live = (uint32_t *)((char *)c->mg_block + l.off_live);
        p.colPitchW = (long long)l.live_words;
        p.V = c->global_rows >= 0 ? c->global_rows : c->V;
        p.regain_rows = 0;
    }
    return p;
}

int build_graph(utmos_ctx *c)
{
    if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
    const SelParams p = make_params(c, false);
    cudaGraph_t graph = nullptr;
    UT_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    int rc = UTMOS_OK;
    for (int i = 0; i < kGraphSteps && rc == UTMOS_OK; ++i) rc = launch_step_pair(c->stream, p, c->n_sms, nullptr);
    cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
    if (rc != UTMOS_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    UT_CUDA(e);
    e = cudaGraphInstantiate(&c->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    UT_CUDA(e);
    return UTMOS_OK;
}

}  // namespace

// ================================================================================================
// extern "C"
// ================================================================================================
extern "C" {

const char *utmos_last_error(void) { return g_error.c_str(); }

const char *utmos_version(void) { return "utmos_b200 0.1.0 (sm_100a)"; }

int utmos_device_count(int *count_out)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    if (count_out) *count_out = n;
    return UTMOS_OK;
}

int utmos_host_alloc(void **ptr_out, int64_t bytes)
{
    if (!ptr_out || bytes < 0) { set_error("host_alloc: bad arguments"); return UTMOS_E_ARG; }
    UT_CUDA(cudaMallocHost(ptr_out, (size_t)std::max<int64_t>(bytes, 16)));
    return UTMOS_OK;
}

int utmos_host_free(void *ptr)
{
    if (ptr) UT_CUDA(cudaFreeHost(ptr));
    return UTMOS_OK;
}

int utmos_create(utmos_ctx **ctx_out, int device, int64_t n_samples, int64_t rows_hint, int af_mode, uint32_t flags)
{
    if (!ctx_out) { set_error("create: ctx_out is null"); return UTMOS_E_ARG; }
    *ctx_out = nullptr;
    if (n_samples <= 0 || n_samples > 0x7fffff00ll) { set_error("create: n_samples out of range"); return UTMOS_E_ARG; }
    if (af_mode < UTMOS_AF_NONE || af_mode > UTMOS_AF_F32) { set_error("create: bad af_mode"); return UTMOS_E_ARG; }
    int n = 0;
    utmos_device_count(&n);
    if (n <= 0) { set_error("no CUDA device visible: utmos_b200 has no CPU fallback"); return UTMOS_E_NOGPU; }
    if (device < 0 || device >= n) { set_error("create: device index out of range"); return UTMOS_E_ARG; }
    Trace tr("create");
    UT_CUDA(cudaSetDevice(device));
    int cc_major = 0, cc_minor = 0, sm_count = 0;
    UT_CUDA(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device));
    UT_CUDA(cudaDeviceGetAttribute(&cc_minor, cudaDevAttrComputeCapabilityMinor, device));
    UT_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
    if (cc_major != 10) {
        set_error(std::string("device is sm_") + std::to_string(cc_major * 10 + cc_minor) +
                  ", this library is built for sm_100a (B200) only");
        return UTMOS_E_NOGPU;
    }
    tr.lap("device attributes");
    utmos_ctx *c = new utmos_ctx();
    c->device = device;
    c->n_sms = sm_count;
    c->S = n_samples;
    c->nW = (int)((n_samples + 31) / 32);
    c->pitchW = (c->nW + 3) / 4 * 4;
    c->af_mode = af_mode;
    c->flags = flags;
    if (flags & UTMOS_F_REF_TIES) {
        // the replay lives in argmax_step_kernel and in the REFT flavour of select_listcluster_kernel: per-step kernels (with
        // a streaming recompute after heavy picks) until the lists can be built, the entry-divided tail from there on.
        // UTMOS_F_STEP_KERNELS given as well: per-step kernels all the way (the round-2 flavour, kept for A/B runs).
        if (af_mode == UTMOS_AF_NONE) c->flags &= ~UTMOS_F_REF_TIES;        // count mode has no float sums to replay
        else if (!(flags & UTMOS_F_STEP_KERNELS)) c->ref_hybrid = true;
    }
    // --af: the 8-CTA owner-computes flavour of the tail until a pick covers fewer than 64 rows (three shared-memory
    // atomics per decrement, two of them 64-bit: measured 10.0 ms against 11.9 ms for the tail of config C3)
    if (af_mode != UTMOS_AF_NONE) c->tail_single_rows = 64;
    int rc = UTMOS_OK;
    do {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
            rc = cuda_fail(cudaGetLastError(), "cudaStreamCreate", __FILE__, __LINE__);
            break;
        }
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;                 // never trim: buffers are re-used across selections
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
        if ((rc = dev_alloc(c, (void **)&c->d_nrows, 16)) != UTMOS_OK) break;
        if (cudaMemsetAsync(c->d_nrows, 0, 16, c->stream) != cudaSuccess) { rc = UTMOS_E_CUDA; break; }
        if ((rc = dev_alloc(c, (void **)&c->d_state, sizeof(SelState))) != UTMOS_OK) break;
        if (cudaMemsetAsync(c->d_state, 0, sizeof(SelState), c->stream) != cudaSuccess) { rc = UTMOS_E_CUDA; break; }
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) { rc = UTMOS_E_CUDA; break; }
        tr.lap("streams + state");
        if (rows_hint > 0) rc = grow_rows(c, rows_hint);
        tr.lap("row buffer");
    } while (0);
    if (rc != UTMOS_OK) { utmos_destroy(c); return rc; }
    *ctx_out = c;
    return UTMOS_OK;
}

int utmos_destroy(utmos_ctx *c)
{
    if (!c) return UTMOS_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    t_resolve(c);
    free_select_state(c);
    ingest_scratch_free(c->scratch, c->stream);
    pinned_release(c);
    for (int i = 0; i < 2; ++i) {
        big_free(c, c->d_stage[i], kStageBytes);
        if (c->d_af_stage[i]) cudaFreeAsync(c->d_af_stage[i], c->stream);
        if (c->h_af_stage[i]) cudaFreeHost(c->h_af_stage[i]);
        if (c->ev_handover[i]) cudaEventDestroy(c->ev_handover[i]);
        if (c->ev_copied[i]) cudaEventDestroy(c->ev_copied[i]);
        if (c->ev_consumed[i]) cudaEventDestroy(c->ev_consumed[i]);
    }
    big_free(c, c->d_rows, (size_t)c->rows_cap * (size_t)c->pitchW * 4);
    if (c->d_af) cudaFreeAsync(c->d_af, c->stream);
    if (c->d_nrows) cudaFreeAsync(c->d_nrows, c->stream);
    if (c->d_state) cudaFreeAsync(c->d_state, c->stream);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    cudaGetLastError();
    delete c;
    return UTMOS_OK;
}

int utmos_append_packed(utmos_ctx *c, const uint8_t *rows, int64_t n_rows, int64_t pitch_bytes, const double *af)
{
    if (!c) { set_error("null context"); return UTMOS_E_ARG; }
    if (pitch_bytes < (c->S + 7) / 8) { set_error("append_packed: pitch smaller than ceil(S/8)"); return UTMOS_E_ARG; }
    UT_CUDA(cudaSetDevice(c->device));
    return append_host(c, RAW_PACKED_MSB, rows, n_rows, pitch_bytes, af);
}

// .jl v2 rows: payload + offsets on the host; decoded on the GPU into the raw staging buffer, then
// ingested exactly like utmos_append_packed rows (same filter, same order).
int utmos_append_packed2(utmos_ctx *c, const uint8_t *payload, const uint64_t *offsets, int64_t n_rows, int idx_bytes,
                         const double *af)
{
    if (!c) { set_error("null context"); return UTMOS_E_ARG; }
    if (n_rows < 0 || (n_rows > 0 && (!payload || !offsets))) { set_error("append_packed2: bad arguments"); return UTMOS_E_ARG; }
    if (idx_bytes != 2 && idx_bytes != 4) { set_error("append_packed2: idx_bytes must be 2 or 4"); return UTMOS_E_ARG; }
    if (idx_bytes == 2 && c->S > 65536) { set_error("append_packed2: 16-bit indices need S <= 65536"); return UTMOS_E_ARG; }
    if (c->finalized) { set_error("append after finalize"); return UTMOS_E_ARG; }
    if (n_rows == 0) return UTMOS_OK;
    const bool want_af = c->af_mode != UTMOS_AF_NONE;
    if (want_af && !af) { set_error("append_packed2: AF required for an AF context"); return UTMOS_E_ARG; }
    UT_CUDA(cudaSetDevice(c->device));
    const long long pitch = (c->S + 7) / 8;
    for (int64_t r = 0; r < n_rows; ++r)
        if (offsets[r + 1] < offsets[r] || offsets[r + 1] - offsets[r] > (uint64_t)pitch) {
            set_error("append_packed2: row offsets must ascend by at most ceil(S/8) bytes");
            return UTMOS_E_ARG;
        }
    UT_TRY(grow_rows(c, c->rows_upper + n_rows));
    const long long chunk_rows = std::max(1ll, (long long)(kStageBytes / (size_t)pitch));
    UT_TRY(ensure_stage(c, want_af ? std::min<long long>(chunk_rows, n_rows) : 0, false));
    uint8_t *d_payload = nullptr;
    unsigned long long *d_off = nullptr;
    int *d_bad = nullptr;
    const size_t pay_cap = (size_t)std::min<long long>(chunk_rows, n_rows) * (size_t)pitch;
    UT_TRY(dev_alloc(c, (void **)&d_payload, pay_cap));
    UT_TRY(dev_alloc(c, (void **)&d_off, ((size_t)std::min<long long>(chunk_rows, n_rows) + 1) * 8));
    UT_TRY(dev_alloc(c, (void **)&d_bad, 4));
    UT_CUDA(cudaMemsetAsync(d_bad, 0, 4, c->stream));
    int rc = UTMOS_OK;
    for (long long r0 = 0; r0 < n_rows && rc == UTMOS_OK; r0 += chunk_rows) {
        const long long n = std::min<long long>(chunk_rows, n_rows - r0);
        const int b = c->stage_next;
        c->stage_next ^= 1;
        if (c->stage_used[b]) UT_CUDA(cudaEventSynchronize(c->ev_consumed[b]));
        const size_t bytes = (size_t)(offsets[r0 + n] - offsets[r0]);
        t_begin(c, T_H2D, c->stream);
        if (bytes) UT_CUDA(cudaMemcpyAsync(d_payload, payload + offsets[r0], bytes, cudaMemcpyHostToDevice, c->stream));
        UT_CUDA(cudaMemcpyAsync(d_off, offsets + r0, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
        const double *d_af_chunk = nullptr;
        if (want_af) {
            UT_CUDA(cudaMemcpyAsync(c->d_af_stage[b], af + r0, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
            d_af_chunk = c->d_af_stage[b];
        }
        t_end(c, c->stream);
        t_begin(c, T_INGEST, c->stream);
        rc = launch_unpack_rows2(c->stream, d_payload, d_off, n, (int)c->S, pitch, idx_bytes, (uint8_t *)c->d_stage[b], d_bad,
                                 &c->n_launch);
        if (rc == UTMOS_OK)
            rc = launch_ingest(c->stream, c->scratch, RAW_PACKED_MSB, c->d_stage[b], n, pitch, d_af_chunk, (int)c->S, c->pitchW,
                               c->d_rows, want_af ? c->d_af : nullptr, c->d_nrows, &c->n_launch);
        t_end(c, c->stream);
        UT_CUDA(cudaEventRecord(c->ev_consumed[b], c->stream));
        c->stage_used[b] = true;
        UT_CUDA(cudaStreamSynchronize(c->stream));            // d_payload / d_off are reused by the next chunk
    }
    int bad = 0;
    if (rc == UTMOS_OK) UT_CUDA(cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost));
    dev_free(c, d_payload, pay_cap);
    dev_free(c, d_off, ((size_t)std::min<long long>(chunk_rows, n_rows) + 1) * 8);
    dev_free(c, d_bad, 4);
    if (rc != UTMOS_OK) return rc;
    c->rows_upper += n_rows;
    if (bad) { set_error("append_packed2: " + std::to_string(bad) + " rows are malformed (index >= S or ragged length)"); return UTMOS_E_DATA; }
    return UTMOS_OK;
}

int utmos_append_packed_device(utmos_ctx *c, const uint8_t *d_rows, int64_t n_rows, int64_t pitch_bytes,
                               const double *d_af)
{
    if (!c) { set_error("null context"); return UTMOS_E_ARG; }
    if (c->finalized) { set_error("append after finalize"); return UTMOS_E_ARG; }
    if (pitch_bytes < (c->S + 7) / 8) { set_error("append_packed: pitch smaller than ceil(S/8)"); return UTMOS_E_ARG; }
    if (n_rows <= 0) return UTMOS_OK;
    const bool want_af = c->af_mode != UTMOS_AF_NONE;
    if (want_af && !d_af) { set_error("append_packed_device: AF required for an AF context"); return UTMOS_E_ARG; }
    UT_CUDA(cudaSetDevice(c->device));
    UT_TRY(grow_rows(c, c->rows_upper + n_rows));
    t_begin(c, T_INGEST, c->stream);
    UT_TRY(launch_ingest(c->stream, c->scratch, RAW_PACKED_MSB, d_rows, n_rows, pitch_bytes, d_af, (int)c->S, c->pitchW,
                         c->d_rows, want_af ? c->d_af : nullptr, c->d_nrows, &c->n_launch));
    t_end(c, c->stream);
    c->rows_upper += n_rows;
    return UTMOS_OK;
}

int utmos_append_dense_u8(utmos_ctx *c, const uint8_t *chunk, int64_t n_rows)
{
    if (!c) { set_error("null context"); return UTMOS_E_ARG; }
    if (c->af_mode != UTMOS_AF_NONE) { set_error("append_dense_u8: bool data needs an af_mode NONE context"); return UTMOS_E_ARG; }
    UT_CUDA(cudaSetDevice(c->device));
    return append_host(c, RAW_DENSE_U8, chunk, n_rows, c->S, nullptr);
}

int utmos_append_dense_f32(utmos_ctx *c, const float *chunk, int64_t n_rows)
{
    if (!c) { set_error("null context"); return UTMOS_E_ARG; }
    if (c->af_mode == UTMOS_AF_NONE) { set_error("append_dense_f32: float data needs an AF context"); return UTMOS_E_ARG; }
    UT_CUDA(cudaSetDevice(c->device));
    return append_host(c, RAW_DENSE_F32, chunk, n_rows, c->S * 4, nullptr);
}

// Opt-in flavour of the chunk streamer for bool data (UTMOS_B200_H5_GPU_LZF=1): the host threads only pread the
// COMPRESSED chunks into pinned staging (header: offsets, lengths, stored-raw flags; then the chunk bytes), one H2D
// copy per batch, and lzf_unpack_bool_kernel decodes and packs them on the GPU.  Every row of an hdf5 file is kept
// (utmos/select.py:153), so the destination of a chunk's rows is known on the host when nothing was appended before.
static int append_h5_chunks_gpu_lzf(utmos_ctx *c, int fd, int64_t n_chunks, const int64_t *addr, const int64_t *nbytes,
                                    const uint32_t *fmask, int64_t rows_per_chunk, int64_t total_rows, int lzf, int nthreads)
{
    UT_TRY(grow_rows(c, total_rows));
    UT_TRY(ensure_stage(c, 0, true));
    int *d_bad = nullptr;
    UT_TRY(dev_alloc(c, (void **)&d_bad, 4));
    UT_CUDA(cudaMemsetAsync(d_bad, 0, 4, c->stream));
    int rc = UTMOS_OK;
    long long rows_left = total_rows, rows_done = 0;
    for (long long c0 = 0; c0 < n_chunks && rc == UTMOS_OK && rows_left > 0;) {
        // chunks of this batch: header + 16-byte aligned chunk bytes must fit one staging buffer
        long long k = 0;
        size_t data = 0;
        while (c0 + k < n_chunks && k < 16384) {
            const size_t sz = ((size_t)nbytes[c0 + k] + 15) / 16 * 16;
            const size_t hdr = ((size_t)(k + 1) * 13 + 15) / 16 * 16;
            if (hdr + data + sz > kStageBytes) break;
            data += sz;
            ++k;
        }
        if (k == 0) { set_error("append_h5_chunks: one compressed chunk exceeds the staging buffer"); rc = UTMOS_E_ARG; break; }
        const size_t hdr = ((size_t)k * 13 + 15) / 16 * 16;
        const int b = c->stage_next;
        c->stage_next ^= 1;
        if (c->stage_used[b]) {
            if (cudaEventSynchronize(c->ev_consumed[b]) != cudaSuccess) { rc = cuda_fail(cudaGetLastError(), "cudaEventSynchronize", __FILE__, __LINE__); break; }
            cudaStreamWaitEvent(c->copy_stream, c->ev_consumed[b], 0);
        }
        uint8_t *base = (uint8_t *)c->h_stage[b];
        long long *h_off = reinterpret_cast<long long *>(base);
        int *h_len = reinterpret_cast<int *>(base + (size_t)k * 8);
        uint8_t *h_raw = base + (size_t)k * 12;
        size_t at = 0;
        for (long long i = 0; i < k; ++i) {
            h_off[i] = (long long)at;
            h_len[i] = (int)nbytes[c0 + i];
            h_raw[i] = (!lzf || (fmask[c0 + i] & 1u)) ? 1 : 0;
            at += ((size_t)nbytes[c0 + i] + 15) / 16 * 16;
        }
        std::atomic<long long> next(0);
        std::atomic<int> bad(0);
        auto worker = [&]() {
            for (;;) {
                const long long i = next.fetch_add(1);
                if (i >= k || bad.load()) return;
                if (pread(fd, base + hdr + (size_t)h_off[i], (size_t)h_len[i], addr[c0 + i]) != (ssize_t)h_len[i]) { bad.store(1); return; }
            }
        };
        {
            std::vector<std::thread> pool;
            const int nt = (int)std::min<long long>(nthreads, k);
            for (int t = 1; t < nt; ++t) pool.emplace_back(worker);
            worker();
            for (auto &t : pool) t.join();
        }
        if (bad.load()) { set_error("append_h5_chunks: short read"); rc = UTMOS_E_DATA; break; }
        const long long n = std::min(rows_left, k * (long long)rows_per_chunk);
        t_begin(c, T_H2D, c->copy_stream);
        if (cudaMemcpyAsync(c->d_stage[b], base, hdr + data, cudaMemcpyHostToDevice, c->copy_stream) != cudaSuccess) { rc = cuda_fail(cudaGetLastError(), "cudaMemcpyAsync", __FILE__, __LINE__); break; }
        t_end(c, c->copy_stream);
        cudaEventRecord(c->ev_copied[b], c->copy_stream);
        cudaStreamWaitEvent(c->stream, c->ev_copied[b], 0);
        const uint8_t *d_base = (const uint8_t *)c->d_stage[b];
        t_begin(c, T_INGEST, c->stream);
        rc = launch_lzf_unpack_bool(c->stream, d_base + hdr, reinterpret_cast<const long long *>(d_base),
                                    reinterpret_cast<const int *>(d_base + (size_t)k * 8), d_base + (size_t)k * 12, k,
                                    (int)rows_per_chunk, n, (int)c->S, c->pitchW, rows_done, c->d_rows, c->d_nrows, d_bad,
                                    &c->n_launch);
        t_end(c, c->stream);
        cudaEventRecord(c->ev_consumed[b], c->stream);
        c->stage_used[b] = true;
        c->rows_upper += n;
        rows_done += n;
        rows_left -= n;
        c0 += k;
    }
    int h_bad = 0;
    if (rc == UTMOS_OK) {
        if (cudaMemcpyAsync(&h_bad, d_bad, 4, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
            cudaStreamSynchronize(c->stream) != cudaSuccess)
            rc = cuda_fail(cudaGetLastError(), "lzf_unpack", __FILE__, __LINE__);
        else if (h_bad) { set_error("append_h5_chunks: malformed LZF chunk"); rc = UTMOS_E_DATA; }
    }
    dev_free(c, d_bad, 4);
    return rc;
}

// hdf5 chunk streamer: the chunks of the 'data' dataset are read (pread) and LZF-decoded by `threads` host threads
// straight into the pinned staging buffers, one staging buffer of chunks at a time; the H2D copy and the packing
// kernels of batch i run on the copy / compute streams while the host decodes batch i+1 into the other buffer.
int utmos_append_h5_chunks(utmos_ctx *c, const char *path, int64_t n_chunks, const int64_t *addr, const int64_t *nbytes,
                           const uint32_t *fmask, int64_t rows_per_chunk, int64_t total_rows, int is_f32, int lzf,
                           int threads)
{
    if (!c || !path || n_chunks < 0 || (n_chunks > 0 && (!addr || !nbytes || !fmask)) || rows_per_chunk <= 0 || total_rows < 0) {
        set_error("append_h5_chunks: bad arguments");
        return UTMOS_E_ARG;
    }
    if (c->finalized) { set_error("append after finalize"); return UTMOS_E_ARG; }
    if ((is_f32 != 0) != (c->af_mode != UTMOS_AF_NONE)) { set_error("append_h5_chunks: data type does not match the context's af_mode"); return UTMOS_E_ARG; }
    if (n_chunks == 0 || total_rows == 0) return UTMOS_OK;
    UT_CUDA(cudaSetDevice(c->device));
    const int kind = is_f32 ? RAW_DENSE_F32 : RAW_DENSE_U8;
    const size_t row_bytes = (size_t)c->S * (is_f32 ? 4 : 1);
    const size_t chunk_bytes = row_bytes * (size_t)rows_per_chunk;
    if (chunk_bytes > kStageBytes) { set_error("append_h5_chunks: one chunk exceeds the staging buffer"); return UTMOS_E_ARG; }
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { set_error(std::string("append_h5_chunks: cannot open ") + path); return UTMOS_E_ARG; }
    // bool chunks are LZF-decoded on the GPU (lzf_unpack_bool_kernel): the host only reads the compressed bytes, so PCIe
    // carries the file size instead of the dense size (C4 at 50,000 x 200,000: 134 ms against 257 ms with the host codec).
    // UTMOS_B200_H5_GPU_LZF=0 keeps the host codec for A/B runs.
    static const bool gpu_lzf = !(getenv("UTMOS_B200_H5_GPU_LZF") && atoi(getenv("UTMOS_B200_H5_GPU_LZF")) == 0);
    if (gpu_lzf && !is_f32 && c->rows_upper == 0 && chunk_bytes / 8 + (8 << 10) + 64 <= (200u << 10)) {
        const int nt = std::max(1, std::min(threads > 0 ? threads : (int)std::thread::hardware_concurrency(), 64));
        const int rc_gpu = append_h5_chunks_gpu_lzf(c, fd, n_chunks, addr, nbytes, fmask, rows_per_chunk, total_rows, lzf, nt);
        close(fd);
        return rc_gpu;
    }
    const long long per_batch = (long long)(kStageBytes / chunk_bytes);
    int rc = grow_rows(c, c->rows_upper + total_rows);
    if (rc == UTMOS_OK) rc = ensure_stage(c, is_f32 ? per_batch * rows_per_chunk : 0, true);
    const int nthreads = std::max(1, std::min(threads > 0 ? threads : (int)std::thread::hardware_concurrency(), 64));
    long long rows_left = total_rows;
    for (long long c0 = 0; c0 < n_chunks && rc == UTMOS_OK && rows_left > 0; c0 += per_batch) {
        const long long k = std::min(per_batch, (long long)n_chunks - c0);
        const int b = c->stage_next;
        c->stage_next ^= 1;
        if (c->stage_used[b]) {
            if (cudaEventSynchronize(c->ev_consumed[b]) != cudaSuccess) { rc = cuda_fail(cudaGetLastError(), "cudaEventSynchronize", __FILE__, __LINE__); break; }
            cudaStreamWaitEvent(c->copy_stream, c->ev_consumed[b], 0);
        }
        uint8_t *dst = (uint8_t *)c->h_stage[b];
        std::atomic<long long> next(0);
        std::atomic<int> bad(0);
        auto worker = [&]() {
            std::vector<uint8_t> raw;
            for (;;) {
                const long long i = next.fetch_add(1);
                if (i >= k || bad.load()) return;
                const long long ci = c0 + i;
                uint8_t *out = dst + (size_t)i * chunk_bytes;
                const bool stored_raw = !lzf || (fmask[ci] & 1u);
                if (stored_raw) {
                    if (nbytes[ci] != (int64_t)chunk_bytes || pread(fd, out, chunk_bytes, addr[ci]) != (ssize_t)chunk_bytes) { bad.store(1); return; }
                } else {
                    raw.resize((size_t)nbytes[ci]);
                    if (pread(fd, raw.data(), raw.size(), addr[ci]) != (ssize_t)raw.size() ||
                        utmos_lzf_decompress(raw.data(), (int64_t)raw.size(), out, (int64_t)chunk_bytes) != (int64_t)chunk_bytes) { bad.store(1); return; }
                }
            }
        };
        {
            std::vector<std::thread> pool;
            const int nt = (int)std::min<long long>(nthreads, k);
            for (int t = 1; t < nt; ++t) pool.emplace_back(worker);
            worker();
            for (auto &t : pool) t.join();
        }
        if (bad.load()) { set_error("append_h5_chunks: short read or malformed LZF chunk"); rc = UTMOS_E_DATA; break; }
        const long long n = std::min(rows_left, k * rows_per_chunk);          // the last chunk is stored full size
        const size_t bytes = (size_t)n * row_bytes;
        t_begin(c, T_H2D, c->copy_stream);
        if (cudaMemcpyAsync(c->d_stage[b], dst, bytes, cudaMemcpyHostToDevice, c->copy_stream) != cudaSuccess) { rc = cuda_fail(cudaGetLastError(), "cudaMemcpyAsync", __FILE__, __LINE__); break; }
        t_end(c, c->copy_stream);
        cudaEventRecord(c->ev_copied[b], c->copy_stream);
        cudaStreamWaitEvent(c->stream, c->ev_copied[b], 0);
        t_begin(c, T_INGEST, c->stream);
        rc = launch_ingest(c->stream, c->scratch, kind, c->d_stage[b], n, (long long)row_bytes, is_f32 ? c->d_af_stage[b] : nullptr,
                           (int)c->S, c->pitchW, c->d_rows, is_f32 ? c->d_af : nullptr, c->d_nrows, &c->n_launch);
        t_end(c, c->stream);
        cudaEventRecord(c->ev_consumed[b], c->stream);
        c->stage_used[b] = true;
        c->rows_upper += n;
        rows_left -= n;
    }
    close(fd);
    return rc;
}

int utmos_finalize(utmos_ctx *c, int64_t *num_vars_out, int64_t *var_count_out)
{
    if (!c) { set_error("null context"); return UTMOS_E_ARG; }
    if (c->finalized) { set_error("finalize called twice"); return UTMOS_E_ARG; }
    Trace tr("finalize");
    UT_CUDA(cudaSetDevice(c->device));
    UT_TRY(sync_all(c));
    tr.lap("wait for ingest");
    long long v = 0;
    UT_CUDA(cudaMemcpy(&v, c->d_nrows, 8, cudaMemcpyDeviceToHost));
    c->V = v;
    if (v >= (1ll << 40)) { set_error("finalize: too many rows"); return UTMOS_E_ARG; }
    const size_t S = (size_t)c->S;
    const long long S32 = (c->S + 31) / 32 * 32;
    c->colPitchW = std::max(8ll, ((v + 31) / 32 + 7) / 8 * 8);
    const bool af = c->af_mode != UTMOS_AF_NONE;
    // limb width: sums of up to V limbs must stay below 2^63 (see oracle_fixed_scale)
    const long long v_all = c->global_rows >= 0 ? std::max(c->global_rows, v) : v;   // every rank must use the same scale
    int lg = 0;
    while ((1ll << lg) < v_all + 1) ++lg;
    c->L = std::min(48, 63 - lg);
    c->scale = 2 * c->L - 1;

    UT_TRY(dev_alloc(c, (void **)&c->d_live, (size_t)c->colPitchW * 4));
    UT_TRY(dev_alloc(c, (void **)&c->d_live0, (size_t)c->colPitchW * 4));
    UT_TRY(dev_alloc(c, (void **)&c->d_gain_cnt, S * 4));
    UT_TRY(dev_alloc(c, (void **)&c->d_gain0_cnt, S * 4));
    UT_TRY(dev_alloc(c, (void **)&c->d_var_count, S * 4));
    UT_TRY(dev_alloc(c, (void **)&c->d_mask, S));
    UT_TRY(dev_alloc(c, (void **)&c->d_weights, S * 8));
    UT_TRY(dev_alloc(c, (void **)&c->d_out_idx, S * 8));
    UT_TRY(dev_alloc(c, (void **)&c->d_out_new, S * 8));
    UT_TRY(dev_alloc(c, (void **)&c->d_out_score, S * 8));
    UT_TRY(dev_alloc(c, (void **)&c->d_out_time, S * 8));
    UT_TRY(dev_alloc(c, (void **)&c->d_dbg, 128));
    for (int i = 0; i < 2; ++i) {
        UT_TRY(dev_alloc(c, (void **)&c->d_list_off[i], S * 4));
        UT_TRY(dev_alloc(c, (void **)&c->d_list_len[i], S * 4));
    }
    UT_TRY(dev_alloc(c, (void **)&c->d_cursor, S * 4));
    UT_TRY(dev_alloc(c, (void **)&c->d_pool_cursor, 16));
    UT_CUDA(cudaMemsetAsync(c->d_dbg, 0, 128, c->stream));
    UT_CUDA(cudaMemsetAsync(c->d_out_time, 0, S * 8, c->stream));
    UT_TRY(dev_alloc(c, (void **)&c->d_dbg_score, S * 8));
    UT_TRY(dev_alloc(c, (void **)&c->d_bar, 64));
    UT_TRY(dev_alloc(c, (void **)&c->d_partials, sizeof(ArgPartial) * 2048));
    UT_CUDA(cudaMemsetAsync(c->d_gain0_cnt, 0, S * 4, c->stream));
    UT_CUDA(cudaMemsetAsync(c->d_var_count, 0, S * 4, c->stream));
    if (af) {
        UT_TRY(dev_alloc(c, (void **)&c->d_gain_lo, S * 8));
        UT_TRY(dev_alloc(c, (void **)&c->d_gain_hi, S * 8));
        UT_TRY(dev_alloc(c, (void **)&c->d_gain0_lo, S * 8));
        UT_TRY(dev_alloc(c, (void **)&c->d_gain0_hi, S * 8));
        UT_TRY(dev_alloc(c, (void **)&c->d_q_lo, (size_t)v * 8));
        UT_TRY(dev_alloc(c, (void **)&c->d_q_hi, (size_t)v * 8));
        UT_CUDA(cudaMemsetAsync(c->d_gain0_lo, 0, S * 8, c->stream));
        UT_CUDA(cudaMemsetAsync(c->d_gain0_hi, 0, S * 8, c->stream));
    }
    tr.lap("small allocations");
    // sample-major copy when it fits (keeps ~2 GiB of head-room)
    if (!(c->flags & UTMOS_F_NO_TRANSPOSE) && v > 0) {
        const size_t bytes = (size_t)S32 * (size_t)c->colPitchW * 4;
        // keep head-room for the edge lists / staging: the copy must leave at least 1/8 of the device free
        int rc_alloc = UTMOS_E_NOMEM;
        size_t free_b = 0, total_b = 0;
        bool roomy = true;
        if (bytes > (8ull << 30)) {                       // only worth asking the driver for multi-GB copies
            UT_CUDA(cudaMemGetInfo(&free_b, &total_b));
            size_t parked = 0;
            for (auto &b : g_big_cache) parked += b.bytes;
            roomy = free_b + parked > bytes + total_b / 8;
        }
        if (roomy) rc_alloc = big_alloc(c, (void **)&c->d_cols, bytes);
        if (rc_alloc != UTMOS_OK) {
            c->d_cols = nullptr;
            if (c->flags & UTMOS_F_FORCE_TRANSPOSE) {
                set_error("finalize: sample-major copy does not fit in device memory");
                return UTMOS_E_NOMEM;
            }
        }
    }
    if ((c->flags & UTMOS_F_REF_TIES) && v > 0 && !c->d_cols) {
        // the replay walks sample-major rows: without that copy (it does not fit, or UTMOS_F_NO_TRANSPOSE) the context falls
        // back to the exact-arithmetic order and says so in utmos_info()[9]; UTMOS_F_FORCE_TRANSPOSE makes it an error
        if (c->flags & UTMOS_F_FORCE_TRANSPOSE) {
            set_error("finalize: UTMOS_F_REF_TIES needs the sample-major copy");
            return UTMOS_E_NOMEM;
        }
        c->flags &= ~(UTMOS_F_REF_TIES | UTMOS_F_STEP_KERNELS);
        c->ref_hybrid = false;
    }
    tr.lap("sample-major allocation");
    if (c->d_cols) {
        t_begin(c, T_TRANSPOSE, c->stream);
        UT_TRY(launch_transpose(c->stream, c->d_rows, v, c->pitchW, (int)c->S, c->d_cols, c->colPitchW, &c->n_launch));
        t_end(c, c->stream);
    }
    t_begin(c, T_GAIN, c->stream);
    if (af) UT_TRY(launch_fixed_af(c->stream, c->d_af, v, c->af_mode, c->L, c->scale, c->d_q_lo, c->d_q_hi, c->d_state,
                                   &c->n_launch));
    UT_TRY(launch_live_init(c->stream, c->d_live0, c->colPitchW, v, c->d_q_lo, c->d_q_hi, af ? 1 : 0, &c->n_launch));
    {
        const SelParams p = make_params(c, true);
        UT_TRY(launch_gain_init(c->stream, p, c->d_var_count, &c->n_launch));
    }
    t_end(c, c->stream);
    tr.lap("launches");
    UT_TRY(sync_all(c));
    tr.lap("transpose + gains (device)");
    SelState st;
    UT_CUDA(cudaMemcpy(&st, c->d_state, sizeof(st), cudaMemcpyDeviceToHost));
    if (st.af_invalid) {
        set_error("finalize: " + std::to_string(st.af_invalid) +
                  " informative rows have an allele frequency that is NaN, negative or > 1");
        return UTMOS_E_DATA;
    }
    c->af_inexact = st.af_inexact;
    if (var_count_out) {
        std::vector<unsigned int> tmp(S);
        UT_CUDA(cudaMemcpy(tmp.data(), c->d_var_count, S * 4, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < S; ++i) var_count_out[i] = tmp[i];
    }
    if (num_vars_out) *num_vars_out = v;
    if (!(c->flags & UTMOS_F_STEP_KERNELS)) UT_TRY(persistent_grid(c->device, &c->grid, &c->block));
    tr.lap("readback + occupancy");
    c->finalized = true;
    return UTMOS_OK;
}

int utmos_select_begin(utmos_ctx *c, const uint8_t *mask, const double *weights)
{
    if (!c || !mask) { set_error("select_begin: null argument"); return UTMOS_E_ARG; }
    if (!c->finalized) { set_error("select_begin before finalize"); return UTMOS_E_ARG; }
    const size_t S = (size_t)c->S;
    for (size_t i = 0; i < S; ++i)
        if (mask[i] > 2) { set_error("select_begin: mask values must be 0, 1 or 2"); return UTMOS_E_ARG; }
    if (weights)
        for (size_t i = 0; i < S; ++i)
            if (!isfinite(weights[i])) { set_error("select_begin: weights must be finite"); return UTMOS_E_ARG; }
    UT_CUDA(cudaSetDevice(c->device));
    UT_CUDA(cudaMemcpyAsync(c->d_mask, mask, S, cudaMemcpyHostToDevice, c->stream));
    std::vector<uint32_t> selw((size_t)c->nW, 0u);
    for (size_t i = 0; i < S; ++i)
        if (mask[i] == 1) selw[i >> 5] |= 1u << (i & 31);
    if (!c->d_selw) UT_TRY(dev_alloc(c, (void **)&c->d_selw, (size_t)c->nW * 4));
    UT_CUDA(cudaMemcpyAsync(c->d_selw, selw.data(), (size_t)c->nW * 4, cudaMemcpyHostToDevice, c->stream));
    const bool had_weights = c->has_weights;
    c->has_weights = weights != nullptr;
    if (weights) UT_CUDA(cudaMemcpyAsync(c->d_weights, weights, S * 8, cudaMemcpyHostToDevice, c->stream));
    UT_CUDA(cudaMemcpyAsync(c->d_live, c->d_live0, (size_t)c->colPitchW * 4, cudaMemcpyDeviceToDevice, c->stream));
    UT_CUDA(cudaMemcpyAsync(c->d_gain_cnt, c->d_gain0_cnt, S * 4, cudaMemcpyDeviceToDevice, c->stream));
    if (c->af_mode != UTMOS_AF_NONE) {
        UT_CUDA(cudaMemcpyAsync(c->d_gain_lo, c->d_gain0_lo, S * 8, cudaMemcpyDeviceToDevice, c->stream));
        UT_CUDA(cudaMemcpyAsync(c->d_gain_hi, c->d_gain0_hi, S * 8, cudaMemcpyDeviceToDevice, c->stream));
    }
    if (c->d_local0_cnt) {
        UT_CUDA(cudaMemcpyAsync(c->d_local_cnt, c->d_local0_cnt, S * 4, cudaMemcpyDeviceToDevice, c->stream));
        if (c->af_mode != UTMOS_AF_NONE) {
            UT_CUDA(cudaMemcpyAsync(c->d_local_lo, c->d_local0_lo, S * 8, cudaMemcpyDeviceToDevice, c->stream));
            UT_CUDA(cudaMemcpyAsync(c->d_local_hi, c->d_local0_hi, S * 8, cudaMemcpyDeviceToDevice, c->stream));
        }
    }
    SelState st, prev;
    UT_CUDA(cudaMemcpyAsync(&prev, c->d_state, sizeof(prev), cudaMemcpyDeviceToHost, c->stream));
    UT_CUDA(cudaStreamSynchronize(c->stream));
    memset(&st, 0, sizeof(st));
    st.winner = -1;
    st.mgpu_seq = prev.mgpu_seq;       // exchange sequence numbers stay monotonic over the life of the context
    UT_CUDA(cudaMemcpyAsync(c->d_state, &st, sizeof(st), cudaMemcpyHostToDevice, c->stream));
    {
        const SelParams p0 = make_params(c, false);
        UT_TRY(launch_sum_gains(c->stream, p0, &c->n_launch));     // st->live_bits = set bits of the scoring rows
    }
    UT_CUDA(cudaStreamSynchronize(c->stream));      // host buffers (mask, weights, st) may go away
    if ((c->flags & UTMOS_F_STEP_KERNELS) && (!c->graph_exec || had_weights != c->has_weights)) UT_TRY(build_graph(c));
    c->lists_valid = false;
    c->lists_cur = 0;                  // the first compaction of a selection goes to the buffer sized for it (no re-allocation)
    c->mg_tail = false;
    c->selecting = true;
    return UTMOS_OK;
}

int utmos_select_steps(utmos_ctx *c, int64_t max_steps, int64_t *idx_out, int64_t *new_out, double *score_out,
                       int64_t *n_out, int *stop_out)
{
    if (!c || !n_out || !stop_out) { set_error("select_steps: null argument"); return UTMOS_E_ARG; }
    if (!c->selecting) { set_error("select_steps before select_begin"); return UTMOS_E_ARG; }
    *n_out = 0;
    *stop_out = UTMOS_STOP_NONE;
    UT_CUDA(cudaSetDevice(c->device));
    SelState st;
    UT_CUDA(cudaMemcpy(&st, c->d_state, sizeof(st), cudaMemcpyDeviceToHost));
    const long long start = st.step;
    if (st.stop != 0) { *stop_out = st.stop; return UTMOS_OK; }
    if (max_steps <= 0) return UTMOS_OK;
    if ((!idx_out || !new_out)) { set_error("select_steps: null output"); return UTMOS_E_ARG; }
    const long long limit = std::min<long long>(c->S, start + max_steps);
    if (limit <= start) {
        // every sample is already used: the reference's next argmax sees only zeros (select.py:43,51)
        st.stop = UTMOS_STOP_ZERO;
        UT_CUDA(cudaMemcpy(c->d_state, &st, sizeof(st), cudaMemcpyHostToDevice));
        *stop_out = UTMOS_STOP_ZERO;
        return UTMOS_OK;
    }
    st.limit = limit;
    UT_CUDA(cudaMemcpy(c->d_state, &st, sizeof(st), cudaMemcpyHostToDevice));
    const SelParams p = make_params(c, false);
    t_begin(c, T_SELECT, c->stream);
    if (c->flags & UTMOS_F_STEP_KERNELS && c->mg_world == 1) {
        while (true) {
            UT_CUDA(cudaGraphLaunch(c->graph_exec, c->stream));
            c->n_launch += 2 * kGraphSteps;
            UT_CUDA(cudaMemcpyAsync(&st, c->d_state, sizeof(st), cudaMemcpyDeviceToHost, c->stream));
            UT_CUDA(cudaStreamSynchronize(c->stream));
            if (st.stop != 0 || st.step >= limit) break;
        }
    } else {
        const bool multi = c->mg_world > 1;
        const bool af = c->af_mode != UTMOS_AF_NONE;
        int CL = 0, tail_ok = 0;
        const bool hybrid = c->ref_hybrid && !multi && p.ref_ties;
        if (!multi && !hybrid && !(c->flags & UTMOS_F_NO_CLUSTER)) UT_TRY(cluster_plan(p, &CL));
        if (!multi && !(c->flags & UTMOS_F_NO_TAIL)) UT_TRY(tail_plan(p, &tail_ok));
        if (hybrid) tail_ok = tail_ok && listcluster_fits(p);      // the only tail flavour with the replay
        if (multi) tail_ok = c->mg_list_cap > 0;          // same decision on every rank (utmos_mgpu_export)
        // heavy picks in count mode: subtract the newly covered rows (cover_decrement_kernel) instead of recomputing
        // every gain from the sample-major copy (regain_kernel).  Default since round 2 (full GPU suite green with it,
        // greedy loop 9.83 -> 9.06 ms on the 1kGP shape); UTMOS_B200_DECREMENT=0 goes back to regain_kernel for A/B runs.
        static const bool use_decrement = !(getenv("UTMOS_B200_DECREMENT") && atoi(getenv("UTMOS_B200_DECREMENT")) == 0);
        // (AF flavours keep regain_kernel: a tiled AF decrement over the newly covered rows was measured on config C3 and
        // bought nothing -- head 4.96 ms against 4.76 ms -- because that head is bound by the cluster kernel's own steps)
        if (use_decrement && !multi && !af && CL > 0 && c->d_cols && !c->d_newmask && !(c->flags & UTMOS_F_STEP_KERNELS))
            UT_TRY(dev_alloc(c, (void **)&c->d_newmask, (size_t)c->colPitchW * 4));
        // Head: greedy steps by the cluster (or grid-wide, or multi-GPU) kernel; a pick that covers very many rows
        // ends the launch and the conditional regain kernel recomputes the gains.  While the tail flavour is still
        // waiting for the live part of the matrix to become sparse, launches are kept short so the host can switch
        // as soon as the per-sample live-row lists fit the budget.
        // Tail: one CTA runs all remaining steps from the lists (on every rank, identically, when multi-GPU).
        const bool wide = c->S > 65535;
        const size_t pool_elem = wide ? 4 : 2;
        const unsigned long long list_budget = multi ? c->mg_list_cap - 64 : (c->list_budget ? c->list_budget : (wide ? kListBudgetWide : kListBudget));
        const size_t estride = af ? 2 : 1;
        c->tail_budget = tail_ok ? list_budget : 0;
        auto reserve_lists = [&](int which, unsigned long long entries) -> int {
            const size_t need = (size_t)std::max<unsigned long long>(entries, 1) * estride;
            if (need > c->lists_cap[which]) {
                if (c->lists_external && which == 0) { set_error("merged edge lists exceed the exchange block"); return UTMOS_E_NOMEM; }
                big_free(c, c->d_lists[which], c->lists_cap[which] * 16);     // stream is idle here (just synchronised)
                UT_TRY(big_alloc(c, (void **)&c->d_lists[which], need * 16));
                c->lists_cap[which] = need;
            }
            return UTMOS_OK;
        };
        MgLayout l = {};
        MgpuParams m;
        memset(&m, 0, sizeof(m));
        if (multi) {
            l = mg_layout((size_t)c->S, c->mg_world, af, c->mg_list_cap, c->mg_merged_rows);
            m.rank = c->mg_rank;
            m.world = c->mg_world;
            m.global_V = c->global_rows >= 0 ? c->global_rows : c->V;
            m.delta_cnt = c->d_delta_cnt;
            m.delta_lo = c->d_delta_lo;
            m.delta_hi = c->d_delta_hi;
            m.local_cnt = c->d_local_cnt;
            m.local_lo = c->d_local_lo;
            m.local_hi = c->d_local_hi;
            if (!m.local_cnt) { set_error("select_steps: utmos_set_gains0 must run before a multi-GPU selection"); return UTMOS_E_ARG; }
            char *mine = (char *)c->mg_block;
            m.inbox_cnt = (unsigned int *)(mine + l.off_cnt);
            m.inbox_lo = (unsigned long long *)(mine + l.off_lo);
            m.inbox_hi = (unsigned long long *)(mine + l.off_hi);
            m.flags = (unsigned long long *)(mine + l.off_flags);
            for (int q = 0; q < c->mg_world; ++q) {
                if (q == c->mg_rank) continue;
                if (!c->mg_peer[q]) { set_error("select_steps: multi-GPU peers are not connected"); return UTMOS_E_ARG; }
                char *pb = (char *)c->mg_peer[q];
                m.peer_inbox_cnt[q] = (unsigned int *)(pb + l.off_cnt);
                m.peer_inbox_lo[q] = (unsigned long long *)(pb + l.off_lo);
                m.peer_inbox_hi[q] = (unsigned long long *)(pb + l.off_hi);
                m.peer_flags[q] = (unsigned long long *)(pb + l.off_flags);
            }
        }
        Trace tr("select_steps");
        auto part_t0 = std::chrono::steady_clock::now();
        auto part_lap = [&](int which) {
            const auto now = std::chrono::steady_clock::now();
            c->part_ms[which] += std::chrono::duration<double, std::milli>(now - part_t0).count();
            part_t0 = now;
        };
        while (true) {
            SelParams q = make_params(c, false);
            // count mode: one SM retires a light pick faster (1.5 us + 3.8 ns per row against 3.5 us flat); AF: two 64-bit
            // shared-memory adds per decrement make one SM the slower choice whatever the pick covers
            // more than 65,535 samples: the alternative is the owner-computes cluster tail, where every CTA walks every entry
            // (50,000 x 2 M: 6.3 against 8.0 us per step; 100,000 x 2 M: 7.4 against 9.3)
            // the same holds for any cohort whose state does not fit one SM (50,000 x 2 M rows: tail 94.8 against 106.6 ms)
            const bool sliced = tail_cluster_size(q, false) != 1;
            const unsigned int heavy_rows = hybrid ? 1u : c->tail_heavy_rows != 0xffffffffu ? c->tail_heavy_rows : ((af || wide || sliced) ? 1u : 768u);
            if (c->lists_valid && heavy_rows > 0 && !(st.tail_single & 2u) && listcluster_fits(q)) {
                // heavy picks: the entries of the pick divided over a 16-CTA cluster, state sliced over its shared memories (select_listcluster_kernel);
                // it hands over (bit 1 of st.tail_single) once a pick covers fewer than tail_heavy_rows rows
                UT_TRY(launch_listcluster(c->stream, q, c->lists_total, heavy_rows, &c->n_launch));
                UT_TRY(launch_sum_gains(c->stream, q, &c->n_launch));
                c->flavour_used = multi ? 5 : 3;
            } else if (c->lists_valid) {
                // heavy picks: cluster of 8 CTAs, each applying the decrements of the samples it owns; light picks: one CTA
                const unsigned int single_rows = (st.tail_single & 1u) ? 0u : c->tail_single_rows;
                const bool cluster = single_rows > 0 || tail_cluster_size(q, false) != 1;     // wide / mid-size cohorts: sliced state
                uint32_t *live_priv = nullptr;
                if (cluster && !tail_live_in_smem(q, cluster)) {
                    const size_t need = (size_t)tail_cluster_size(q, cluster) * (size_t)q.colPitchW * 4;
                    if (need > c->live_priv_bytes) {
                        dev_free(c, c->d_live_priv, c->live_priv_bytes);
                        UT_TRY(dev_alloc(c, (void **)&c->d_live_priv, need));
                        c->live_priv_bytes = need;
                    }
                    live_priv = c->d_live_priv;
                }
                UT_TRY(launch_tail(c->stream, q, c->lists_total, cluster, single_rows, live_priv, &c->n_launch));
                UT_TRY(launch_sum_gains(c->stream, q, &c->n_launch));
                c->flavour_used = multi ? 5 : 3;
            } else if (multi) {
                m.seq0 = st.mgpu_seq;
                UT_TRY(launch_mgpu(c->stream, q, m, c->mg_grid, c->mg_block_threads, c->d_bar, c->d_partials, &c->n_launch));
                UT_TRY(launch_sum_gains(c->stream, q, &c->n_launch));
                c->flavour_used = 4;
            } else if (hybrid) {
                // reference tie order, head: per-step kernels (argmax with the replay, cover; a pick that covers >= regain_rows
                // rows only clears live bits and leaves st->regain set for the streaming recompute), a few steps per host check
                for (int rep = 0; rep < 4; ++rep) {
                    UT_TRY(launch_step_pair(c->stream, q, c->n_sms, &c->n_launch));
                    UT_TRY(launch_regain(c->stream, q, &c->n_launch));
                }
                UT_TRY(launch_sum_gains(c->stream, q, &c->n_launch));
                c->flavour_used = 0;
            } else {
                // a few launches are queued between host checks; once the live part is sparse enough the queued
                // head kernels return immediately (they test st->live_bits at launch)
                static const int head_reps = getenv("UTMOS_B200_HEAD_REPS") ? std::max(1, std::min(64, atoi(getenv("UTMOS_B200_HEAD_REPS")))) : 4;
                for (int rep = 0; rep < head_reps; ++rep) {
                    if (CL > 0) {
                        UT_TRY(launch_cluster(c->stream, q, CL, &c->n_launch, c->d_newmask));
                        c->flavour_used = 2;
                        c->cluster = CL;
                    } else {
                        UT_TRY(launch_persistent(c->stream, q, c->grid, c->block, c->d_bar, c->d_partials, &c->n_launch));
                        c->flavour_used = 1;
                    }
                    if (!q.regain_rows) {
                        // no recompute configured (no sample-major copy): still refresh st->live_bits, the host decides
                        // the hand-over to the tail from it
                        UT_TRY(launch_sum_gains(c->stream, q, &c->n_launch));
                        break;
                    }
                    if (CL > 0 && c->d_newmask) UT_TRY(launch_cover_decrement(c->stream, q, c->d_newmask, &c->n_launch));
                    else UT_TRY(launch_regain(c->stream, q, &c->n_launch));
                    UT_TRY(launch_sum_gains(c->stream, q, &c->n_launch));
                }
            }
            UT_CUDA(cudaMemcpyAsync(&st, c->d_state, sizeof(st), cudaMemcpyDeviceToHost, c->stream));
            UT_CUDA(cudaStreamSynchronize(c->stream));
            tr.lap(c->lists_valid ? "tail launch" : "head launches");
            part_lap(c->lists_valid ? P_TAIL : P_HEAD);
            if (st.stop != 0 || st.step >= limit || st.abort_flag) break;
            if (hybrid && c->lists_valid && st.tie_step) {
                // a tie candidate with more live rows than the tail's sort buffer: this one step by the per-step kernels
                UT_CUDA(cudaMemsetAsync(&c->d_state->tie_step, 0, sizeof(unsigned int), c->stream));
                UT_TRY(launch_step_pair(c->stream, q, c->n_sms, &c->n_launch));
                UT_TRY(launch_regain(c->stream, q, &c->n_launch));
                UT_TRY(launch_sum_gains(c->stream, q, &c->n_launch));
                st.tie_step = 0;
                continue;
            }
            if (hybrid && !c->lists_valid && tail_ok && st.step > 0) {
                // the step kernels do not look at the hand-over rule: the host does, from the last pick's new_count
                long long last_new = 0;
                UT_CUDA(cudaMemcpy(&last_new, c->d_out_new + (st.step - 1), sizeof(last_new), cudaMemcpyDeviceToHost));
                st.want_tail = last_new < (long long)q.tail_rows ? 1u : 0u;
            }
            if (tail_ok && !c->lists_valid && st.want_tail && st.live_bits <= list_budget) {
                if (multi) {
                    // replicated tail: every rank writes its live rows into the merged lists of all ranks (mgpu.cu)
                    GatherParams g;
                    memset(&g, 0, sizeof(g));
                    g.rank = c->mg_rank; g.world = c->mg_world; g.S = (int)c->S;
                    g.lcnt = c->d_lcnt;
                    g.inbox_cnt = m.inbox_cnt;
                    g.flags = m.flags;
                    g.st = c->d_state;
                    g.live_word0 = c->mg_row_base / 32;
                    g.live_words = (c->V + 31) / 32;
                    EdgeDst d;
                    memset(&d, 0, sizeof(d));
                    d.world = 1;                      // entries are built locally, whole segments are pushed afterwards
                    for (int r = 0; r < c->mg_world; ++r) {
                        char *blk = r == c->mg_rank ? (char *)c->mg_block : (char *)c->mg_peer[r];
                        g.peer_inbox_cnt[r] = (unsigned int *)(blk + l.off_cnt);
                        g.peer_flags[r] = (unsigned long long *)(blk + l.off_flags);
                        g.live_dst[r] = (uint32_t *)(blk + l.off_live);
                        g.lists_dst[r] = (uint4 *)(blk + l.off_lists);
                        g.pool_dst[r] = (unsigned char *)(blk + l.off_pool);
                    }
                    char *mine = (char *)c->mg_block;
                    d.lists[0] = (uint4 *)(mine + l.off_lists);
                    d.pool[0] = (unsigned short *)(mine + l.off_pool);
                    d.slot_base = c->d_my_base;
                    d.pool_base = c->d_pool_base;
                    d.row_base = c->mg_row_base;
                    g.my_base = c->d_my_base;
                    g.pool_base = c->d_pool_base;
                    g.pool_cursor = c->d_pool_cursor;
                    g.estride = (int)estride;
                    g.pool_elem = (int)pool_elem;
                    // my merged live mask: zero the padding words; the ranks fill their own word ranges
                    UT_CUDA(cudaMemsetAsync(mine + l.off_live, 0, l.live_words * 4, c->stream));
                    UT_TRY(launch_live_counts(c->stream, q, c->d_lcnt, &c->n_launch));
                    g.seq = st.mgpu_seq + 1;       // (a) all-gather of the per-rank live counts; also orders the memset above
                    UT_TRY(launch_gather_counts(c->stream, g, &c->n_launch));
                    UT_TRY(launch_gather_offsets(c->stream, g, c->d_list_off[0], c->d_list_len[0], c->d_my_base, c->d_cursor,
                                                 c->d_pool_base, &c->n_launch));
                    UT_TRY(launch_build_edges(c->stream, q, d, c->d_cursor, c->d_pool_cursor, &c->n_launch));
                    UT_TRY(launch_gather_push(c->stream, g, &c->n_launch));
                    UT_TRY(launch_gather_live(c->stream, g, q.live, &c->n_launch));
                    g.seq = st.mgpu_seq + 2;       // (b) everybody's entries have landed everywhere
                    UT_TRY(launch_gather_done(c->stream, g, &c->n_launch));
                    st.mgpu_seq += 2;
                    UT_CUDA(cudaMemcpyAsync(&c->d_state->mgpu_seq, &st.mgpu_seq, sizeof(st.mgpu_seq), cudaMemcpyHostToDevice, c->stream));
                    UT_CUDA(cudaStreamSynchronize(c->stream));
                    part_lap(P_HANDOVER);
                    SelState chk;
                    UT_CUDA(cudaMemcpy(&chk, c->d_state, sizeof(chk), cudaMemcpyDeviceToHost));
                    if (chk.abort_flag) { st.abort_flag = chk.abort_flag; break; }
                    if (!c->lists_external) {
                        big_free(c, c->d_lists[0], c->lists_cap[0] * 16);
                        big_free(c, c->d_pool, c->pool_cap * pool_elem);
                    }
                    c->d_lists[0] = (uint4 *)(mine + l.off_lists);
                    c->lists_cap[0] = (size_t)c->mg_list_cap * estride;
                    c->d_pool = (unsigned short *)(mine + l.off_pool);
                    c->pool_cap = l.pool_cap;
                    c->lists_external = true;
                    c->lists_cur = 0;
                    c->mg_tail = true;
                    c->lists_total = st.live_bits;
                    c->lists_valid = true;
                    continue;
                }
                // first compaction: edge lists from the bit matrix
                if (!c->ev_handover[0]) {
                    UT_CUDA(cudaEventCreate(&c->ev_handover[0]));
                    UT_CUDA(cudaEventCreate(&c->ev_handover[1]));
                }
                UT_CUDA(cudaEventRecord(c->ev_handover[0], c->stream));
                UT_TRY(reserve_lists(c->lists_cur, st.live_bits));
                // pooled carrier lists are padded to 8 entries per row (rows with >= 7 carriers): <= 15/7 per live bit
                const size_t pool_need = (size_t)mgpu_pool_share(st.live_bits);
                if (pool_need > c->pool_cap) {
                    big_free(c, c->d_pool, c->pool_cap * pool_elem);
                    c->pool_cap = pool_need;
                    UT_TRY(big_alloc(c, (void **)&c->d_pool, c->pool_cap * pool_elem));
                }
                q = make_params(c, false);
                UT_TRY(launch_build_lists(c->stream, q, c->d_lists[c->lists_cur], c->d_list_off[c->lists_cur],
                                          c->d_list_len[c->lists_cur], c->d_cursor, c->d_pool, c->d_pool_cursor,
                                          &c->n_launch));
                UT_CUDA(cudaEventRecord(c->ev_handover[1], c->stream));
                c->handover_unread = true;
                c->lists_total = st.live_bits;
                c->lists_valid = true;
                tr.lap("list allocation + build launch");
            } else if (c->lists_valid && st.recompact) {
                // re-compaction: keep the entries whose row is still live
                const int nxt = c->lists_cur ^ 1;
                UT_TRY(reserve_lists(nxt, st.live_bits));
                q = make_params(c, false);
                UT_TRY(launch_filter_lists(c->stream, q, c->d_lists[c->lists_cur], c->d_list_off[c->lists_cur],
                                           c->d_list_len[c->lists_cur], c->d_lists[nxt], c->d_list_off[nxt],
                                           c->d_list_len[nxt], &c->n_launch));
                c->lists_cur = nxt;
                c->lists_total = st.live_bits;
                tr.lap("re-compaction launch");
            }
        }
    }
    t_end(c, c->stream);
    UT_TRY(sync_all(c));
    if (c->ev_handover[0] && cudaEventQuery(c->ev_handover[1]) == cudaSuccess && c->handover_unread) {
        float ho = 0.f;
        if (cudaEventElapsedTime(&ho, c->ev_handover[0], c->ev_handover[1]) == cudaSuccess) {
            c->part_ms[P_HANDOVER] += ho;          // it was lapped as part of the first tail launch
            c->part_ms[P_TAIL] -= ho;
        }
        c->handover_unread = false;
    }
    UT_CUDA(cudaMemcpy(&st, c->d_state, sizeof(st), cudaMemcpyDeviceToHost));
    if (st.abort_flag) {
        set_error("select_steps: device watchdog tripped (grid barrier timeout)");
        return UTMOS_E_DEVICE;
    }
    const long long n = st.step - start;
    if (n > 0) {
        UT_CUDA(cudaMemcpy(idx_out, c->d_out_idx + start, (size_t)n * 8, cudaMemcpyDeviceToHost));
        UT_CUDA(cudaMemcpy(new_out, c->d_out_new + start, (size_t)n * 8, cudaMemcpyDeviceToHost));
        if (score_out) UT_CUDA(cudaMemcpy(score_out, c->d_out_score + start, (size_t)n * 8, cudaMemcpyDeviceToHost));
    }
    *n_out = n;
    *stop_out = st.stop;
    return UTMOS_OK;
}

int utmos_debug_gains(utmos_ctx *c, int64_t *count_out, double *score_out)
{
    if (!c || !c->selecting) { set_error("debug_gains: no selection in progress"); return UTMOS_E_ARG; }
    UT_CUDA(cudaSetDevice(c->device));
    const size_t S = (size_t)c->S;
    const SelParams p = make_params(c, false);
    if (count_out) {
        std::vector<unsigned int> tmp(S);
        UT_CUDA(cudaMemcpy(tmp.data(), c->d_gain_cnt, S * 4, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < S; ++i) count_out[i] = tmp[i];
    }
    if (score_out) {
        UT_TRY(launch_debug_scores(c->stream, p, c->d_dbg_score, &c->n_launch));
        UT_CUDA(cudaStreamSynchronize(c->stream));
        UT_CUDA(cudaMemcpy(score_out, c->d_dbg_score, S * 8, cudaMemcpyDeviceToHost));
    }
    return UTMOS_OK;
}


// ---- resume -------------------------------------------------------------------------------
// The state of a selection in progress is small: the working sample mask (picked samples are 0), the live-row mask
// and the report rows emitted so far.  Gains are not part of it -- utmos_select_import recomputes them from the matrix
// and the live mask (one streaming pass), so a checkpoint can never hold gains that disagree with its masks.

int utmos_select_export(utmos_ctx *c, uint8_t *mask_out, uint32_t *live_out, int64_t live_words, int64_t *idx_out,
                        int64_t *new_out, double *score_out, int64_t rows_cap, int64_t *n_rows_out, int64_t *tot_out,
                        int *stop_out)
{
    if (!c || !c->selecting || !mask_out || !live_out || !n_rows_out) { set_error("select_export: no selection in progress / null argument"); return UTMOS_E_ARG; }
    if (c->mg_world > 1) { set_error("select_export: not available for a row-sharded (multi-GPU) matrix"); return UTMOS_E_ARG; }
    if (live_words != c->colPitchW) { set_error("select_export: live mask must have utmos_info()[8] words"); return UTMOS_E_ARG; }
    UT_CUDA(cudaSetDevice(c->device));
    UT_CUDA(cudaStreamSynchronize(c->stream));
    SelState st;
    UT_CUDA(cudaMemcpy(&st, c->d_state, sizeof(st), cudaMemcpyDeviceToHost));
    if (st.step > rows_cap) { set_error("select_export: row buffers too small"); return UTMOS_E_ARG; }
    UT_CUDA(cudaMemcpy(mask_out, c->d_mask, (size_t)c->S, cudaMemcpyDeviceToHost));
    UT_CUDA(cudaMemcpy(live_out, c->d_live, (size_t)c->colPitchW * 4, cudaMemcpyDeviceToHost));
    if (st.step > 0) {
        if (idx_out) UT_CUDA(cudaMemcpy(idx_out, c->d_out_idx, (size_t)st.step * 8, cudaMemcpyDeviceToHost));
        if (new_out) UT_CUDA(cudaMemcpy(new_out, c->d_out_new, (size_t)st.step * 8, cudaMemcpyDeviceToHost));
        if (score_out) UT_CUDA(cudaMemcpy(score_out, c->d_out_score, (size_t)st.step * 8, cudaMemcpyDeviceToHost));
    }
    *n_rows_out = st.step;
    if (tot_out) *tot_out = st.tot;
    if (stop_out) *stop_out = st.stop;
    return UTMOS_OK;
}

int utmos_select_import(utmos_ctx *c, const uint8_t *mask, const double *weights, const uint32_t *live, int64_t live_words,
                        const int64_t *idx, const int64_t *new_, const double *score, int64_t n_rows, int64_t tot, int stop)
{
    if (!c || !mask || !live || n_rows < 0 || (n_rows > 0 && (!idx || !new_))) { set_error("select_import: bad arguments"); return UTMOS_E_ARG; }
    if (!c->finalized) { set_error("select_import before finalize"); return UTMOS_E_ARG; }
    if (c->mg_world > 1) { set_error("select_import: not available for a row-sharded (multi-GPU) matrix"); return UTMOS_E_ARG; }
    if (live_words != c->colPitchW || n_rows > c->S) { set_error("select_import: state does not fit this matrix"); return UTMOS_E_ARG; }
    UT_TRY(utmos_select_begin(c, mask, weights));            // validates mask / weights, resets the selection state
    const size_t S = (size_t)c->S;
    UT_CUDA(cudaMemcpyAsync(c->d_live, live, (size_t)c->colPitchW * 4, cudaMemcpyHostToDevice, c->stream));
    if (n_rows > 0) {
        UT_CUDA(cudaMemcpyAsync(c->d_out_idx, idx, (size_t)n_rows * 8, cudaMemcpyHostToDevice, c->stream));
        UT_CUDA(cudaMemcpyAsync(c->d_out_new, new_, (size_t)n_rows * 8, cudaMemcpyHostToDevice, c->stream));
        if (score) UT_CUDA(cudaMemcpyAsync(c->d_out_score, score, (size_t)n_rows * 8, cudaMemcpyHostToDevice, c->stream));
    }
    // gains of the restored live mask: popcount (and AF limbs) of col & live, or one reduction per set bit of the live rows
    UT_CUDA(cudaMemsetAsync(c->d_gain_cnt, 0, S * 4, c->stream));
    const bool af = c->af_mode != UTMOS_AF_NONE;
    if (af) {
        UT_CUDA(cudaMemsetAsync(c->d_gain_lo, 0, S * 8, c->stream));
        UT_CUDA(cudaMemsetAsync(c->d_gain_hi, 0, S * 8, c->stream));
    }
    {
        const SelParams p = make_params(c, false);
        UT_TRY(launch_gain_init(c->stream, p, nullptr, &c->n_launch));
    }
    SelState st;
    UT_CUDA(cudaMemcpyAsync(&st, c->d_state, sizeof(st), cudaMemcpyDeviceToHost, c->stream));
    UT_CUDA(cudaStreamSynchronize(c->stream));
    st.step = n_rows;
    st.tot = tot;
    st.stop = stop;
    UT_CUDA(cudaMemcpyAsync(c->d_state, &st, sizeof(st), cudaMemcpyHostToDevice, c->stream));
    {
        const SelParams p = make_params(c, false);
        UT_TRY(launch_sum_gains(c->stream, p, &c->n_launch));
    }
    UT_CUDA(cudaStreamSynchronize(c->stream));
    return UTMOS_OK;
}

// ---- multi-GPU plumbing ------------------------------------------------------------------------------------

int utmos_rows(utmos_ctx *c, int64_t *rows_out)
{
    if (!c || !rows_out) { set_error("rows: null argument"); return UTMOS_E_ARG; }
    UT_CUDA(cudaSetDevice(c->device));
    if (c->finalized) { *rows_out = c->V; return UTMOS_OK; }
    UT_TRY(sync_all(c));
    long long v = 0;
    UT_CUDA(cudaMemcpy(&v, c->d_nrows, 8, cudaMemcpyDeviceToHost));
    *rows_out = v;
    return UTMOS_OK;
}

int utmos_mgpu_layout(utmos_ctx *c, int64_t row_base, int64_t merged_rows, int allow_tail)
{
    if (!c || row_base < 0 || merged_rows < row_base || (row_base & 31) || (merged_rows & 31)) { set_error("mgpu_layout: bad arguments"); return UTMOS_E_ARG; }
    if (c->mg_block) { set_error("mgpu_layout after mgpu_export"); return UTMOS_E_ARG; }
    c->mg_row_base = row_base;
    c->mg_merged_rows = merged_rows;
    c->mg_allow_tail = allow_tail != 0;
    return UTMOS_OK;
}

int utmos_mgpu_export(utmos_ctx *c, int rank, int world, uint8_t *handle_out)
{
    if (!c || !handle_out || world < 1 || world > kMaxRanks || rank < 0 || rank >= world) { set_error("mgpu_export: bad arguments"); return UTMOS_E_ARG; }
    if (!c->finalized) { set_error("mgpu_export before finalize"); return UTMOS_E_ARG; }
    if (c->mg_block) { set_error("mgpu_export called twice"); return UTMOS_E_ARG; }
    UT_CUDA(cudaSetDevice(c->device));
    const bool af = c->af_mode != UTMOS_AF_NONE;
    const size_t S = (size_t)c->S;
    // merged edge lists for the replicated tail: as many entries as the tail hand-over budget allows (utmos_set_gains0
    // has stored the set bits of all ranks' scoring rows); every rank computes the same capacity
    c->mg_list_cap = 0;
    if (c->mg_allow_tail && c->mg_merged_rows > 0 && c->mg_merged_rows < 0xffffffffll && tail_possible(c->S, af ? 1 : 0) &&
        !(c->flags & (UTMOS_F_NO_TAIL | UTMOS_F_NO_TRANSPOSE)))
        c->mg_list_cap = std::min<unsigned long long>(c->list_budget ? c->list_budget : (c->S > 65535 ? kListBudgetWide : kListBudget), c->total_bits) + 64;
    const MgLayout l = mg_layout(S, world, af, c->mg_list_cap, c->mg_merged_rows);
    cudaIpcMemHandle_t h;
    UT_TRY(mg_block_acquire(c->device, l.bytes, &c->mg_block, &h));
    UT_CUDA(cudaMemsetAsync(c->mg_block, 0, l.small_bytes, c->stream));     // inboxes and flags; big regions are rewritten before use
    c->mg_bytes = l.bytes;
    c->mg_rank = rank;
    c->mg_world = world;
    UT_TRY(dev_alloc(c, (void **)&c->d_delta_cnt, S * 4));
    UT_CUDA(cudaMemsetAsync(c->d_delta_cnt, 0, S * 4, c->stream));
    if (af) {
        UT_TRY(dev_alloc(c, (void **)&c->d_delta_lo, S * 8));
        UT_TRY(dev_alloc(c, (void **)&c->d_delta_hi, S * 8));
        UT_CUDA(cudaMemsetAsync(c->d_delta_lo, 0, S * 8, c->stream));
        UT_CUDA(cudaMemsetAsync(c->d_delta_hi, 0, S * 8, c->stream));
    }
    UT_TRY(dev_alloc(c, (void **)&c->d_lcnt, S * 4));
    UT_TRY(dev_alloc(c, (void **)&c->d_my_base, S * 4));
    UT_TRY(dev_alloc(c, (void **)&c->d_pool_base, 16));
    UT_CUDA(cudaStreamSynchronize(c->stream));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle_out, &h, 64);
    UT_TRY(mgpu_grid(c->device, &c->mg_grid, &c->mg_block_threads));
    return UTMOS_OK;
}

int utmos_mgpu_connect(utmos_ctx *c, const uint8_t *handles)
{
    if (!c || !handles || !c->mg_block) { set_error("mgpu_connect: export first"); return UTMOS_E_ARG; }
    UT_CUDA(cudaSetDevice(c->device));
    for (int q = 0; q < c->mg_world; ++q) {
        if (q == c->mg_rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)q * 64, 64);
        UT_TRY(mg_peer_map(c->device, h, &c->mg_peer[q]));
    }
    return UTMOS_OK;
}

int utmos_get_gains0(utmos_ctx *c, uint32_t *cnt_out, uint64_t *lo_out, uint64_t *hi_out)
{
    if (!c || !c->finalized || !cnt_out) { set_error("get_gains0: bad arguments"); return UTMOS_E_ARG; }
    UT_CUDA(cudaSetDevice(c->device));
    const size_t S = (size_t)c->S;
    UT_CUDA(cudaMemcpy(cnt_out, c->d_gain0_cnt, S * 4, cudaMemcpyDeviceToHost));
    if (c->af_mode != UTMOS_AF_NONE && lo_out && hi_out) {
        UT_CUDA(cudaMemcpy(lo_out, c->d_gain0_lo, S * 8, cudaMemcpyDeviceToHost));
        UT_CUDA(cudaMemcpy(hi_out, c->d_gain0_hi, S * 8, cudaMemcpyDeviceToHost));
    }
    return UTMOS_OK;
}

int utmos_set_gains0(utmos_ctx *c, const uint32_t *cnt, const uint64_t *lo, const uint64_t *hi, int64_t global_rows)
{
    if (!c || !c->finalized || !cnt) { set_error("set_gains0: bad arguments"); return UTMOS_E_ARG; }
    UT_CUDA(cudaSetDevice(c->device));
    const size_t S = (size_t)c->S;
    const bool af = c->af_mode != UTMOS_AF_NONE;
    if (!c->d_local0_cnt) {
        // keep THIS rank's share of the step-0 gains: the multi-GPU head recomputes shares after heavy picks
        UT_TRY(dev_alloc(c, (void **)&c->d_local0_cnt, S * 4));
        UT_TRY(dev_alloc(c, (void **)&c->d_local_cnt, S * 4));
        if (af) {
            UT_TRY(dev_alloc(c, (void **)&c->d_local0_lo, S * 8));
            UT_TRY(dev_alloc(c, (void **)&c->d_local0_hi, S * 8));
            UT_TRY(dev_alloc(c, (void **)&c->d_local_lo, S * 8));
            UT_TRY(dev_alloc(c, (void **)&c->d_local_hi, S * 8));
        }
        UT_CUDA(cudaMemcpyAsync(c->d_local0_cnt, c->d_gain0_cnt, S * 4, cudaMemcpyDeviceToDevice, c->stream));
        if (af) {
            UT_CUDA(cudaMemcpyAsync(c->d_local0_lo, c->d_gain0_lo, S * 8, cudaMemcpyDeviceToDevice, c->stream));
            UT_CUDA(cudaMemcpyAsync(c->d_local0_hi, c->d_gain0_hi, S * 8, cudaMemcpyDeviceToDevice, c->stream));
        }
        UT_CUDA(cudaStreamSynchronize(c->stream));
    }
    UT_CUDA(cudaMemcpy(c->d_gain0_cnt, cnt, S * 4, cudaMemcpyHostToDevice));
    if (c->af_mode != UTMOS_AF_NONE && lo && hi) {
        UT_CUDA(cudaMemcpy(c->d_gain0_lo, lo, S * 8, cudaMemcpyHostToDevice));
        UT_CUDA(cudaMemcpy(c->d_gain0_hi, hi, S * 8, cudaMemcpyHostToDevice));
    }
    if (global_rows >= 0) c->global_rows = global_rows;
    c->total_bits = 0;
    for (size_t i = 0; i < S; ++i) c->total_bits += cnt[i];
    return UTMOS_OK;
}

int utmos_debug_step_times(utmos_ctx *c, int64_t first, int64_t n, int64_t *ns_out)
{
    if (!c || !c->selecting || !ns_out || first < 0 || n < 0 || first + n > c->S) { set_error("debug_step_times: bad arguments"); return UTMOS_E_ARG; }
    UT_CUDA(cudaSetDevice(c->device));
    if (n > 0) UT_CUDA(cudaMemcpy(ns_out, c->d_out_time + first, (size_t)n * 8, cudaMemcpyDeviceToHost));
    return UTMOS_OK;
}

int utmos_debug_counters(utmos_ctx *c, int64_t *out16)
{
    if (!c || !c->finalized || !out16) { set_error("debug_counters: bad arguments"); return UTMOS_E_ARG; }
    UT_CUDA(cudaSetDevice(c->device));
    UT_CUDA(cudaMemcpy(out16, c->d_dbg, 128, cudaMemcpyDeviceToHost));
    return UTMOS_OK;
}

int utmos_set_option(utmos_ctx *c, int option, int64_t value)
{
    if (!c) { set_error("set_option: null context"); return UTMOS_E_ARG; }
    if (option == UTMOS_OPT_REGAIN_ROWS) { c->regain_rows = value; return UTMOS_OK; }
    if (option == UTMOS_OPT_GLOBAL_ROWS) { c->global_rows = value; return UTMOS_OK; }
    if (option == UTMOS_OPT_STEP_TIMES) { c->dbg_time = value != 0; return UTMOS_OK; }
    if (option == UTMOS_OPT_TAIL_ROWS) { c->tail_rows = (unsigned int)std::max<int64_t>(0, value); return UTMOS_OK; }
    if (option == UTMOS_OPT_TIE_ROW_CAP) { c->tie_row_cap = (unsigned int)std::max<int64_t>(0, std::min<int64_t>(value, 8192)); return UTMOS_OK; }
    if (option == UTMOS_OPT_LIST_BUDGET) {
        if (c->finalized) { set_error("set_option: the list budget is fixed at finalize"); return UTMOS_E_ARG; }
        c->list_budget = (unsigned long long)std::max<int64_t>(0, std::min<int64_t>(value, 1ll << 30));
        return UTMOS_OK;
    }
    if (option == UTMOS_OPT_TAIL_HEAVY_ROWS) { c->tail_heavy_rows = value < 0 ? 0xffffffffu : (unsigned int)std::min<int64_t>(value, 0x7fffffff); return UTMOS_OK; }
    if (option == UTMOS_OPT_TAIL_SINGLE_ROWS) { c->tail_single_rows = (unsigned int)std::max<int64_t>(0, value); return UTMOS_OK; }
    set_error("set_option: unknown option");
    return UTMOS_E_ARG;
}

int utmos_info(utmos_ctx *c, int64_t *info, int n)
{
    if (!c || !info) { set_error("info: null argument"); return UTMOS_E_ARG; }
    const int64_t vals[10] = {c->V, (int64_t)c->pitchW * 4, c->d_cols ? 1 : 0, (int64_t)c->dev_bytes, c->scale,
                              (int64_t)c->af_inexact, c->n_launch, c->flavour_used, (int64_t)c->colPitchW,
                              ((c->flags & UTMOS_F_REF_TIES) && c->mg_world <= 1) ? 1 : 0};
    for (int i = 0; i < n && i < 10; ++i) info[i] = vals[i];
    return UTMOS_OK;
}

int utmos_timings(utmos_ctx *c, double *ms, int n, int reset)
{
    if (!c) { set_error("timings: null context"); return UTMOS_E_ARG; }
    for (int i = 0; i < n && i < T_COUNT + P_COUNT; ++i)
        if (ms) ms[i] = i < T_COUNT ? c->ms[i] : c->part_ms[i - T_COUNT];
    if (reset) {
        for (int i = 0; i < T_COUNT; ++i) c->ms[i] = 0.0;
        for (int i = 0; i < P_COUNT; ++i) c->part_ms[i] = 0.0;
    }
    return UTMOS_OK;
}

// Device-side stopwatch for callers that time whole selections (bench.py): both ends synchronise the device and
// record a CUDA event, so the elapsed time covers every kernel and copy of every context in between.
static cudaEvent_t g_timer_ev[kMaxRanks * 2] = {nullptr};

int utmos_timer_start(int device)
{
    if (device < 0 || device >= kMaxRanks) { set_error("timer: device index out of range"); return UTMOS_E_ARG; }
    UT_CUDA(cudaSetDevice(device));
    for (int i = 0; i < 2; ++i)
        if (!g_timer_ev[device * 2 + i]) UT_CUDA(cudaEventCreate(&g_timer_ev[device * 2 + i]));
    UT_CUDA(cudaDeviceSynchronize());
    UT_CUDA(cudaEventRecord(g_timer_ev[device * 2], 0));
    return UTMOS_OK;
}

int utmos_timer_stop(int device, double *ms_out)
{
    if (device < 0 || device >= kMaxRanks || !ms_out || !g_timer_ev[device * 2]) { set_error("timer: not started"); return UTMOS_E_ARG; }
    UT_CUDA(cudaSetDevice(device));
    UT_CUDA(cudaDeviceSynchronize());
    UT_CUDA(cudaEventRecord(g_timer_ev[device * 2 + 1], 0));
    UT_CUDA(cudaEventSynchronize(g_timer_ev[device * 2 + 1]));
    float ms = 0.f;
    UT_CUDA(cudaEventElapsedTime(&ms, g_timer_ev[device * 2], g_timer_ev[device * 2 + 1]));
    *ms_out = ms;
    return UTMOS_OK;
}

static double g_convert_kernel_ms = 0.0;

int utmos_convert_kernel_ms(double *ms_out)
{
    if (!ms_out) { set_error("convert_kernel_ms: null argument"); return UTMOS_E_ARG; }
    *ms_out = g_convert_kernel_ms;
    return UTMOS_OK;
}

int utmos_convert_gt(int device, const int8_t *gt, int64_t n_vars, int64_t n_samples, int64_t ploidy,
                     uint8_t *packed_out, double *af_out, int64_t *num_het_out, int64_t *num_hom_out,
                     uint8_t *singleton_out)
{
    return utmos_convert_gt_ex(device, gt, n_vars, n_samples, ploidy, packed_out, af_out, num_het_out, num_hom_out,
                               singleton_out, 0);
}

int utmos_convert_gt_ex(int device, const int8_t *gt, int64_t n_vars, int64_t n_samples, int64_t ploidy,
                        uint8_t *packed_out, double *af_out, int64_t *num_het_out, int64_t *num_hom_out,
                        uint8_t *singleton_out, int flags)
{
    if (n_vars < 0 || n_samples <= 0 || ploidy <= 0 || ploidy > 8) { set_error("convert_gt: bad shape"); return UTMOS_E_ARG; }
    if (n_vars > 0 && (!gt || !packed_out || !af_out)) { set_error("convert_gt: null buffer"); return UTMOS_E_ARG; }
    int n = 0;
    utmos_device_count(&n);
    if (n <= 0) { set_error("no CUDA device visible: utmos_b200 has no CPU fallback"); return UTMOS_E_NOGPU; }
    if (device < 0 || device >= n) { set_error("convert_gt: device index out of range"); return UTMOS_E_ARG; }
    UT_CUDA(cudaSetDevice(device));
    const long long pitch = (n_samples + 7) / 8;
    const size_t row_in = (size_t)n_samples * (size_t)ploidy;
    const long long chunk = std::max<long long>(1, std::min<long long>(n_vars, (long long)((256ull << 20) / row_in)));
    int8_t *d_gt = nullptr;
    uint8_t *d_packed = nullptr, *d_single = nullptr;
    double *d_af = nullptr;
    unsigned long long *d_hh = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int rc = UTMOS_OK, launches = 0;
    unsigned long long hh[2] = {0, 0};
    do {
        if (n_vars == 0) break;
#define CV(call) if ((call) != cudaSuccess) { rc = cuda_fail(cudaGetLastError(), #call, __FILE__, __LINE__); break; }
        CV(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        CV(cudaMalloc(&d_gt, (size_t)chunk * row_in));
        CV(cudaMalloc(&d_packed, (size_t)chunk * (size_t)pitch));
        CV(cudaMalloc(&d_single, (size_t)chunk));
        CV(cudaMalloc(&d_af, (size_t)chunk * 8));
        CV(cudaMalloc(&d_hh, 16));
        CV(cudaMemsetAsync(d_hh, 0, 16, stream));
        g_convert_kernel_ms = 0.0;
        CV(cudaEventCreate(&ev0));
        CV(cudaEventCreate(&ev1));
        for (long long r0 = 0; r0 < n_vars && rc == UTMOS_OK; r0 += chunk) {
            const long long m = std::min(chunk, n_vars - r0);
            CV(cudaMemcpyAsync(d_gt, gt + (size_t)r0 * row_in, (size_t)m * row_in, cudaMemcpyHostToDevice, stream));
            CV(cudaEventRecord(ev0, stream));
            rc = launch_convert_gt(stream, d_gt, m, (int)n_samples, (int)ploidy, d_packed, pitch, d_af, d_hh, d_single,
                                   (flags & UTMOS_CVT_DROP_SINGLETONS) ? 1 : 0, &launches);
            if (rc != UTMOS_OK) break;
            CV(cudaEventRecord(ev1, stream));
            CV(cudaMemcpyAsync(packed_out + (size_t)r0 * (size_t)pitch, d_packed, (size_t)m * (size_t)pitch,
                               cudaMemcpyDeviceToHost, stream));
            CV(cudaMemcpyAsync(af_out + r0, d_af, (size_t)m * 8, cudaMemcpyDeviceToHost, stream));
            if (singleton_out) CV(cudaMemcpyAsync(singleton_out + r0, d_single, (size_t)m, cudaMemcpyDeviceToHost, stream));
            CV(cudaStreamSynchronize(stream));
            float kms = 0.f;
            if (cudaEventElapsedTime(&kms, ev0, ev1) == cudaSuccess) g_convert_kernel_ms += kms;
        }
        if (rc != UTMOS_OK) break;
        CV(cudaMemcpy(hh, d_hh, 16, cudaMemcpyDeviceToHost));
#undef CV
    } while (0);
    if (d_gt) cudaFree(d_gt);
    if (d_packed) cudaFree(d_packed);
    if (d_single) cudaFree(d_single);
    if (d_af) cudaFree(d_af);
    if (d_hh) cudaFree(d_hh);
    if (stream) cudaStreamDestroy(stream);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (rc != UTMOS_OK) return rc;
    if (num_het_out) *num_het_out = (int64_t)hh[0];
    if (num_hom_out) *num_hom_out = (int64_t)hh[1];
    return UTMOS_OK;
}

}  // extern "C"
