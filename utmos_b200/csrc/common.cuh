// common.cuh -- shared device helpers and host-side declarations for libutmos_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "utmos_b200.h"

namespace utmos {

// ------------------------------------------------------------------------------------------------
// host-side error plumbing
// ------------------------------------------------------------------------------------------------
void set_error(const std::string &msg);
int cuda_fail(cudaError_t err, const char *what, const char *file, int line);

#define UT_CUDA(call)                                                         \
    do {                                                                      \
        cudaError_t ut_err__ = (call);                                        \
        if (ut_err__ != cudaSuccess) return ::utmos::cuda_fail(ut_err__, #call, __FILE__, __LINE__); \
    } while (0)

#define UT_TRY(call)                     \
    do {                                 \
        int ut_rc__ = (call);            \
        if (ut_rc__ != UTMOS_OK) return ut_rc__; \
    } while (0)

// ------------------------------------------------------------------------------------------------
// device state shared by the selection kernels
// ------------------------------------------------------------------------------------------------
struct SelState {
    long long step;        // report rows emitted so far in this selection
    long long limit;       // step budget of the current utmos_select_steps call (absolute)
    long long tot;         // tot_captured (utmos/select.py:99)
    int stop;              // UTMOS_STOP_*
    int winner;            // sample picked by the last argmax (step kernels), -1 = none
    unsigned int abort_flag;   // set by the grid-barrier watchdog
    unsigned int af_inexact;   // AF values with bits below 2^-scale (informative rows only)
    unsigned int af_invalid;   // AF NaN / negative / > 1 on an informative row
    unsigned int want_tail;    // head kernel returned because the tail kernel should take over
    unsigned int recompact;    // tail kernel returned because at most half of its list entries are still live
    unsigned int regain;       // 1 = gains are stale: the pick retired too many rows to subtract, recompute them
    unsigned int tail_single;  // bit 0: the owner-computes cluster tail handed over to the single-CTA flavour; bit 1: the
                               // entry-divided cluster tail handed over to the shared-memory flavours (picks got light)
    unsigned int tie_step;     // reference tie order: the tail met a candidate it cannot replay; run this step with the step kernels
    unsigned int pad0;
    unsigned long long mgpu_seq;    // multi-GPU exchange sequence number (monotonic over the selection)
    unsigned long long live_bits;   // sum of all gains = set bits in rows not yet covered (sum_gains_kernel)
};

struct ArgPartial {        // one per CTA of the persistent / cluster kernel
    double score;
    int idx;
    unsigned int cnt;
    unsigned long long sum;    // sum of the gains this CTA owns (live set bits), for the hand-over to the tail kernel
};

struct SelParams {
    const uint32_t *rows;              // variant-major bit matrix [V][pitchW], sample s = word s>>5, bit s&31
    const uint32_t *cols;              // sample-major copy [S32][colPitchW], row r = word r>>5, bit r&31 (may be null)
    uint32_t *live;                    // [colPitchW] bit r set = row r not yet covered (and scoring)
    unsigned int *gain_cnt;            // [S] rows still uncovered that carry sample s  (new_count if picked now)
    unsigned long long *gain_lo;       // [S] AF flavours: sum of low limbs of AF*2^scale over those rows
    unsigned long long *gain_hi;       // [S] ... high limbs
    const unsigned long long *q_lo;    // [V] per-row fixed-point AF, low limb (L bits)
    const unsigned long long *q_hi;    // [V] high limb
    uint8_t *mask;                     // [S] working copy of sample_mask
    const uint32_t *selw;              // [nW] bit s set = sample s was selectable (mask == 1) at select_begin
    const double *weights;             // [S] or null
    const double *af_vals;             // [V] per-row AF as given (UTMOS_F_REF_TIES replays the reference's float64 sums from it)
    int ref_ties;                      // 1: order exact-arithmetic near-ties like the reference does (step-kernel flavour)
    unsigned int tie_row_cap;          // REFT tail: rows of a candidate the replay may sort (0 = what fits, at most 8,192)
    int af_f32;                        // 1: the rows are float32 GT*AF (hdf5 flavour): AF rounds through float first
    long long *out_idx;                // [S] report rows
    long long *out_new;
    double *out_score;
    long long *out_time;               // [S] %globaltimer (ns) when the pick of each step was made
    SelState *st;
    const uint4 *lists;                // per-sample edge lists (tail kernel), may be null; see tail.cu
    const unsigned int *list_off;      // [S] first entry of sample s in lists
    const unsigned int *list_len;      // [S] entries of sample s
    const unsigned short *pool;        // carrier lists of rows with more than six carriers
    long long *dbg;                    // [16] profiling counters (clock64 cycles per phase, summed over steps)
    long long V;                       // informative rows == num_vars
    long long colPitchW;               // words per sample-major row (multiple of 8)
    int S;
    int pitchW;                        // words per variant-major row (multiple of 4)
    int nW;                            // ceil(S/32)
    int L;                             // limb bits
    int scale;                         // fixed-point scale: value = AF * 2^scale
    int af;                            // 0 count mode, 1 AF flavours
    unsigned long long tail_budget;    // head kernels hand over to the tail kernel once st->live_bits <= this (0 = never)
    unsigned int tail_rows;            // ... and the pick at hand newly covers fewer rows than this
    int dbg_time;                      // 1: record %globaltimer per pick (profiling; costs latency)
    int dsmem_gains;                   // 1: cluster kernel keeps the gains in distributed shared memory
    unsigned int regain_rows;          // picks that newly cover >= this many rows trigger a gain recompute (0 = never)
};

// Multi-GPU exchange (variants sharded by rows, gains replicated on every rank).  All pointers are valid on THIS
// device; peer_* point into the other ranks' memory (CUDA IPC, NVLink P2P).
constexpr int kMaxRanks = 8;
struct MgpuParams {
    int rank, world;
    long long global_V;                       // num_vars summed over the ranks (tot_captured stop rule)
    unsigned long long seq0;                  // exchange sequence number before the first step of this launch
    unsigned int *delta_cnt;                  // [S] this rank's decrements of the current step (two's complement)
    unsigned long long *delta_lo, *delta_hi;  // [S] AF limbs
    unsigned int *local_cnt;                  // [S] THIS rank's share of the gains (live rows of its shard per sample)
    unsigned long long *local_lo, *local_hi;  //     a heavy pick recomputes them by streaming: delta = new - old
    unsigned int *inbox_cnt;                  // [2][world][S] written by the peers (double buffered by step parity)
    unsigned long long *inbox_lo, *inbox_hi;
    unsigned long long *flags;                // [world] last sequence number each peer has published here
    unsigned int *peer_inbox_cnt[kMaxRanks];
    unsigned long long *peer_inbox_lo[kMaxRanks], *peer_inbox_hi[kMaxRanks];
    unsigned long long *peer_flags[kMaxRanks];
};

// Where build_edges_kernel writes the per-sample edge lists (tail.cu).  Single GPU: world = 1 and the context's own
// buffers.  Multi-GPU: the same slots of every rank's merged buffers (replicated tail, see mgpu.cu).
struct EdgeDst {
    int world;
    uint4 *lists[kMaxRanks];
    unsigned short *pool[kMaxRanks];
    const unsigned int *slot_base;     // [S] first entry slot of sample s for THIS rank's rows
    const unsigned int *pool_base;     // device scalar: first pool slot of this rank (null = 0)
    long long row_base;                // added to local row ids (merged live mask numbering)
};

// Multi-GPU hand-over to the replicated tail: all-gather of the per-rank live counts and completion flags.
struct GatherParams {
    int rank, world, S;
    unsigned long long seq;                    // sequence number of this exchange
    const unsigned int *lcnt;                  // [S] live rows of THIS rank that carry sample s
    unsigned int *inbox_cnt;                   // [2][world][S] mine (peers write their slot)
    unsigned long long *flags;                 // [world] mine
    unsigned int *peer_inbox_cnt[kMaxRanks];
    unsigned long long *peer_flags[kMaxRanks];
    uint32_t *live_dst[kMaxRanks];             // merged live masks of every rank
    uint4 *lists_dst[kMaxRanks];               // merged edge lists of every rank (own entry = local pointer)
    unsigned char *pool_dst[kMaxRanks];        // merged carrier pools of every rank (bytes)
    const unsigned int *my_base;               // [S] first merged slot of this rank's entries of sample s
    const unsigned int *pool_base;             // device scalar: first pool element of this rank
    const unsigned int *pool_cursor;           // device scalar: pool elements this rank used (after build_edges_kernel)
    int estride;                               // uint4 per entry (1, or 2 for AF flavours)
    int pool_elem;                             // bytes per pool element (2, or 4 for wide cohorts)
    long long live_word0;                      // first word of this rank's rows in the merged mask
    long long live_words;                      // words this rank contributes (ceil(V/32))
    SelState *st;
};

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ long long global_timer_ns()
{
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// 128-bit streaming load that does not allocate in L1 (data is touched once)
__device__ __forceinline__ uint4 ld_stream_u128(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// .jl bytes are MSB-first (np.packbits); the device layout is LSB-first 32-bit words.
// le = little-endian load of 4 consecutive bytes b0..b3; result bit i <-> sample 32w+i.
__device__ __forceinline__ uint32_t msb_bytes_to_word(uint32_t le) { return __brev(__byte_perm(le, 0, 0x0123)); }
__device__ __forceinline__ uint32_t word_to_msb_bytes(uint32_t w) { return __byte_perm(__brev(w), 0, 0x0123); }

// value = hi * 2^L + lo as an unsigned 128-bit integer, rounded ONCE to the nearest double (ties to even)
// and scaled by 2^-scale.  Bit-for-bit the same function as u128_to_double_scaled in oracle/greedy_oracle.c.
__device__ __forceinline__ double fixed_to_double(unsigned long long lo, unsigned long long hi, int L, int scale)
{
    unsigned long long x_lo = lo + (hi << L);
    unsigned long long x_hi = (hi >> (64 - L)) + (x_lo < lo ? 1ull : 0ull);
    if ((x_lo | x_hi) == 0ull) return 0.0;
    const int msb = x_hi ? 127 - __clzll((long long)x_hi) : 63 - __clzll((long long)x_lo);
    if (msb <= 52) return scalbn((double)x_lo, -scale);
    const int shift = msb - 52;                  // 1..75 low bits are dropped
    unsigned long long mant, rem_hi, rem_lo, half_hi, half_lo;
    if (shift < 64) {
        mant = (x_lo >> shift) | (x_hi ? (x_hi << (64 - shift)) : 0ull);
        rem_lo = x_lo & ((1ull << shift) - 1ull);
        rem_hi = 0ull;
        half_lo = 1ull << (shift - 1);
        half_hi = 0ull;
    } else {
        const int s2 = shift - 64;
        mant = x_hi >> s2;
        rem_lo = x_lo;
        rem_hi = s2 ? (x_hi & ((1ull << s2) - 1ull)) : 0ull;
        if (s2 == 0) { half_hi = 0ull; half_lo = 1ull << 63; }
        else { half_hi = 1ull << (s2 - 1); half_lo = 0ull; }
    }
    const bool above = rem_hi > half_hi || (rem_hi == half_hi && rem_lo > half_lo);
    const bool tie = rem_hi == half_hi && rem_lo == half_lo;
    if (above || (tie && (mant & 1ull))) mant += 1ull;
    return scalbn((double)mant, shift - scale);
}

// np.argmax order: larger score wins, equal scores -> lower index (utmos/select.py:48)
__device__ __forceinline__ bool arg_better(double sa, int ia, double sb, int ib)
{
    return sa > sb || (sa == sb && ia < ib);
}

#endif  // __CUDACC__

// ------------------------------------------------------------------------------------------------
// host launchers implemented in the .cu files
// ------------------------------------------------------------------------------------------------
enum RawKind { RAW_PACKED_MSB = 0, RAW_DENSE_U8 = 1, RAW_DENSE_F32 = 2 };

// ingest.cu : raw chunk (device) -> informative rows appended to the variant-major matrix
struct IngestScratch {
    uint8_t *flags = nullptr;          // [chunk rows]
    unsigned int *block_counts = nullptr;
    unsigned int *block_offsets = nullptr;
    unsigned long long *tile_state = nullptr;   // [0] ticket, [1 + t] decoupled look-back state of tile t
    long long cap_rows = 0;
};
int ingest_scratch_reserve(IngestScratch &sc, long long rows, cudaStream_t stream);
void ingest_scratch_free(IngestScratch &sc, cudaStream_t stream);
int launch_ingest(cudaStream_t stream, IngestScratch &sc, int kind, const void *raw, long long n_rows,
                  long long pitch_in, const double *af_in, int S, int pitchW, uint32_t *rows_out, double *af_out,
                  long long *d_nrows, int *n_launch);

int launch_unpack_rows2(cudaStream_t stream, const uint8_t *payload, const unsigned long long *off, long long n_rows, int S,
                        long long pitch, int idx_bytes, uint8_t *raw_out, int *bad, int *n_launch);

int launch_lzf_unpack_bool(cudaStream_t stream, const uint8_t *blob, const long long *off, const int *len,
                           const uint8_t *stored_raw, long long n_chunks, int rows_per_chunk, long long rows_in_batch,
                           int S, int pitchW, long long row0, uint32_t *rows_out, long long *d_nrows, int *bad,
                           int *n_launch);

// select.cu
int launch_transpose(cudaStream_t stream, const uint32_t *rows, long long V, int pitchW, int S, uint32_t *cols,
                     long long colPitchW, int *n_launch);
int launch_fixed_af(cudaStream_t stream, const double *af, long long V, int af_mode, int L, int scale,
                    unsigned long long *q_lo, unsigned long long *q_hi, SelState *st, int *n_launch);
int launch_live_init(cudaStream_t stream, uint32_t *live, long long colPitchW, long long V,
                     const unsigned long long *q_lo, const unsigned long long *q_hi, int af, int *n_launch);
int launch_gain_init(cudaStream_t stream, const SelParams &p, unsigned int *var_count, int *n_launch);
int launch_step_pair(cudaStream_t stream, const SelParams &p, int n_sms, int *n_launch);
int persistent_grid(int device, int *grid_out, int *block_out);
int launch_persistent(cudaStream_t stream, const SelParams &p, int grid, int block, unsigned int *bar_counter,
                      ArgPartial *partials, int *n_launch);
int launch_debug_scores(cudaStream_t stream, const SelParams &p, double *score_out, int *n_launch);
int launch_regain(cudaStream_t stream, const SelParams &p, int *n_launch);
int launch_cover_decrement(cudaStream_t stream, const SelParams &p, const uint32_t *newmask, int *n_launch);
int launch_sum_gains(cudaStream_t stream, const SelParams &p, int *n_launch);
int launch_build_lists(cudaStream_t stream, const SelParams &p, uint4 *lists, unsigned int *list_off,
                       unsigned int *list_len, unsigned int *cursor, unsigned short *pool, unsigned int *pool_cursor,
                       int *n_launch);
int launch_filter_lists(cudaStream_t stream, const SelParams &p, const uint4 *old_lists, const unsigned int *old_off,
                        const unsigned int *old_len, uint4 *new_lists, unsigned int *new_off, unsigned int *new_len,
                        int *n_launch);
int launch_build_edges(cudaStream_t stream, const SelParams &p, const EdgeDst &d, unsigned int *cursor,
                       unsigned int *pool_cursor, int *n_launch);
int tail_plan(const SelParams &p, int *ok_out);
// mgpu.cu
int launch_live_counts(cudaStream_t stream, const SelParams &p, unsigned int *lcnt, int *n_launch);
int launch_gather_counts(cudaStream_t stream, const GatherParams &g, int *n_launch);
int launch_gather_offsets(cudaStream_t stream, const GatherParams &g, unsigned int *list_off, unsigned int *list_len,
                          unsigned int *my_base, unsigned int *cursor, unsigned int *pool_base, int *n_launch);
int launch_gather_live(cudaStream_t stream, const GatherParams &g, const uint32_t *live, int *n_launch);
int launch_gather_push(cudaStream_t stream, const GatherParams &g, int *n_launch);
int launch_gather_done(cudaStream_t stream, const GatherParams &g, int *n_launch);
unsigned long long mgpu_pool_share(unsigned long long live_bits);
int launch_tail(cudaStream_t stream, const SelParams &p, unsigned long long lists_total, bool cluster,
                unsigned int single_rows, uint32_t *live_priv, int *n_launch);
bool listcluster_fits(const SelParams &p);
int launch_listcluster(cudaStream_t stream, const SelParams &p, unsigned long long lists_total, unsigned int light_rows,
                       int *n_launch);
int tail_live_in_smem(const SelParams &p, bool cluster);
int tail_cluster_size(const SelParams &p, bool cluster);
int tail_possible(long long S, int af);
int launch_mgpu(cudaStream_t stream, const SelParams &p, const MgpuParams &m, int grid, int block,
                unsigned int *bar_counter, ArgPartial *partials, int *n_launch);
int mgpu_grid(int device, int *grid_out, int *block_out);
int cluster_plan(const SelParams &p, int *cluster_out);
int launch_cluster(cudaStream_t stream, const SelParams &p, int CL, int *n_launch, uint32_t *newmask = nullptr);

// convert.cu
int launch_convert_gt(cudaStream_t stream, const int8_t *gt, long long V, int S, int ploidy, uint8_t *packed,
                      long long pitch_out, double *af, unsigned long long *het_hom, uint8_t *singleton,
                      int drop_single, int *n_launch);

}  // namespace utmos
