/*
 * utmos_b200.h -- C ABI of the B200-native greedy maximum-coverage selection path of utmos.
 *
 * One shared library (utmos_b200/libutmos_b200.so, CUDA sm_100a) exports exactly these symbols.  The host
 * side (Python, ctypes: utmos_b200/_native.py) is the only caller.  Plain pointers and sizes only; the
 * caller owns every host buffer, the library owns every device buffer until utmos_destroy().  Every
 * function returns 0 on success or a negative UTMOS_E* code; utmos_last_error() returns the message of the
 * last failure on the calling thread.  One host thread per context; all CUDA streams are internal.
 *
 * The reference has no FFI: the boundary this ABI replaces is the set of NumPy call sites below
 * (file:line into the reference tree).
 *
 *   utmos_append_packed        utmos/select.py:272-280   np.unpackbits + .any(axis=1) filter of one .jl part
 *   utmos_append_dense_u8/f32  utmos/select.py:37        row iteration of the hdf5 'data' dataset
 *   utmos_append_h5_chunks     utmos/select.py:250-251, :37   the same, chunks read + LZF-decoded natively
 *                              utmos/select.py:219-231   (bool, or float32 GT*AF flavour)
 *   utmos_finalize             utmos/select.py:281-284   var_count column sums; :314-320 concat / AF matrix
 *   utmos_select_begin         utmos/select.py:168-187   sample_mask / sample_weights hand-over
 *   utmos_select_steps         utmos/select.py:24-53     calculate_scores (scores, mask, weights, argmax)
 *                              utmos/select.py:91-112    greedy_select loop and its three stop rules
 *   utmos_convert_gt           utmos/convert.py:57-87    is_het|is_hom_alt, het/hom totals, max-alt AF, packbits
 *   utmos_vcf_parse_gt (+ gz)  utmos/convert.py:50-53    allel.read_vcf genotype parse (host threads, BGZF aware)
 */
#ifndef UTMOS_B200_H
#define UTMOS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct utmos_ctx utmos_ctx;

/* error codes */
#define UTMOS_OK 0
#define UTMOS_E_CUDA (-1)     /* a CUDA runtime call failed (message has the CUDA error string) */
#define UTMOS_E_ARG (-2)      /* invalid argument / call order */
#define UTMOS_E_NOGPU (-3)    /* no usable CUDA device: there is no CPU fallback */
#define UTMOS_E_DATA (-4)     /* input data the path cannot represent (e.g. NaN / negative AF on a used row) */
#define UTMOS_E_NOMEM (-5)    /* device memory exhausted */
#define UTMOS_E_DEVICE (-6)   /* device-side watchdog tripped (grid barrier timeout) */

/* AF flavour of a context (utmos/select.py:317-320 float64, :218-223 float32 hdf5 flavour) */
#define UTMOS_AF_NONE 0
#define UTMOS_AF_F64 1
#define UTMOS_AF_F32 2

/* utmos_create flags */
#define UTMOS_F_NO_TRANSPOSE 1u      /* never build the sample-major copy (probe the variant-major matrix) */
#define UTMOS_F_STEP_KERNELS 2u      /* per-step kernel launches (CUDA graph) instead of the persistent kernel */
#define UTMOS_F_FORCE_TRANSPOSE 4u   /* fail instead of falling back when the sample-major copy does not fit */
#define UTMOS_F_NO_CLUSTER 8u        /* grid-wide persistent kernel instead of the one-cluster DSMEM kernel */
#define UTMOS_F_NO_TAIL 16u          /* never switch to the single-CTA list-driven tail kernel */
#define UTMOS_F_REF_TIES 64u         /* --af: order exact-arithmetic (near-)ties by replaying the reference's sequential float64 sums
                                      * (utmos/select.py:37-48): per-step kernels for the head, the entry-divided cluster tail with the replay
                                      * inside from the hand-over on; with UTMOS_F_STEP_KERNELS: per-step kernels all the way.  Needs the
                                      * sample-major copy, one GPU */
#define UTMOS_F_DSMEM_GAINS 32u      /* cluster kernel keeps the gains in distributed shared memory (default: L2 atomics) */

/* stop reasons reported by utmos_select_steps (utmos/select.py:91, :51-52/:93-96, :110-112) */
#define UTMOS_STOP_NONE 0            /* max_steps of this call done; selection can continue */
#define UTMOS_STOP_ZERO 1            /* best masked+weighted score == 0: "Ran out of new variants (multi-allelics)" */
#define UTMOS_STOP_ALL 2             /* tot_captured >= num_vars after the emitted row: "Ran out of new variants" */

const char *utmos_last_error(void);
const char *utmos_version(void);
int utmos_device_count(int *count_out);

/* Pinned host memory for callers that want zero-staging H2D copies (numpy arrays can wrap it). */
int utmos_host_alloc(void **ptr_out, int64_t bytes);
int utmos_host_free(void *ptr);

/*
 * Create a selection context on CUDA device `device` for `n_samples` samples.  `rows_hint` (>= 0) sizes
 * the first device allocation; appends beyond it grow the buffer.
 */
int utmos_create(utmos_ctx **ctx_out, int device, int64_t n_samples, int64_t rows_hint, int af_mode,
                 uint32_t flags);
int utmos_destroy(utmos_ctx *ctx);

/*
 * ".jl v2" rows (the "both axis pack" of the reference's README.md:59-63, which the reference never implemented;
 * SURVEY.md section 8 "next" item 3).  Row v occupies payload[offsets[v] .. offsets[v+1]): either its ceil(S/8)
 * np.packbits bytes, or -- when shorter -- the little-endian uint16 (idx_bytes 2) / uint32 (idx_bytes 4) sample indices of
 * its carriers.  Decoded on the GPU, then filtered and stored exactly like utmos_append_packed rows.
 */
int utmos_append_packed2(utmos_ctx *ctx, const uint8_t *payload, const uint64_t *offsets, int64_t n_rows, int idx_bytes,
                         const double *af);

/*
 * Append one .jl part: `n_rows` rows of `pitch_bytes` (>= ceil(S/8)) bytes, MSB-first bit packed
 * (np.packbits, utmos/convert.py:85), with one float64 AF per row (may be NULL when af_mode == NONE).
 * Rows without any of the first S bits set are dropped together with their AF (utmos/select.py:276-280).
 * `rows` / `af` are HOST pointers (pageable or pinned).  The _device variant takes DEVICE pointers on
 * the context's device (used by the benchmark's resident-input timing).
 */
int utmos_append_packed(utmos_ctx *ctx, const uint8_t *rows, int64_t n_rows, int64_t pitch_bytes,
                        const double *af);
int utmos_append_packed_device(utmos_ctx *ctx, const uint8_t *d_rows, int64_t n_rows, int64_t pitch_bytes,
                               const double *d_af);

/*
 * Append a dense hdf5 chunk: n_rows x S bytes (bool, nonzero = present) or n_rows x S float32 holding
 * GT*AF (utmos/select.py:219-231); for the float flavour the per-row AF is recovered from the row.
 * Host pointers; chunks are streamed through pinned staging on a side stream.  Every row is kept (the
 * file's row count is num_vars, select.py:153; the uninformative-row filter ran before the file was written).
 */
int utmos_append_dense_u8(utmos_ctx *ctx, const uint8_t *chunk, int64_t n_rows);
int utmos_append_dense_f32(utmos_ctx *ctx, const float *chunk, int64_t n_rows);

/*
 * hdf5 chunk streamer (utmos/select.py:250-251 re-using a --lowmem file; :37 iterating its rows): the host side
 * (utmos_b200/h5lite.py) walks the chunk B-tree and passes the byte range of every chunk of the 'data' dataset in
 * row order; the library reads them (pread), LZF-decodes them with `threads` host threads (0 = all cores) straight
 * into its pinned staging buffers and overlaps decode, H2D copy and the packing kernels.  rows_per_chunk x S
 * elements per chunk (bool bytes, or float32 GT*AF when is_f32); total_rows trims the last, partial chunk;
 * lzf = 0 when the dataset has no filter; fmask bit 0 = this chunk is stored raw.
 */
int utmos_append_h5_chunks(utmos_ctx *ctx, const char *path, int64_t n_chunks, const int64_t *addr,
                           const int64_t *nbytes, const uint32_t *fmask, int64_t rows_per_chunk, int64_t total_rows,
                           int is_f32, int lzf, int threads);

/*
 * Close ingestion: builds the sample-major copy (unless it does not fit / is disabled), the per-sample
 * totals and the initial gains.  num_vars_out = informative rows (utmos/select.py:153, data.shape[0]);
 * var_count_out[S] = per-sample totals over all informative rows (utmos/select.py:281-284).
 */
int utmos_finalize(utmos_ctx *ctx, int64_t *num_vars_out, int64_t *var_count_out);

/*
 * (Re)start a selection: mask[S] (1 selectable, 2 excluded, 0 already used; utmos/select.py:168-175) and
 * weights[S] or NULL (utmos/select.py:181-187).  Resets coverage, so one context can run many selections.
 */
int utmos_select_begin(utmos_ctx *ctx, const uint8_t *mask, const double *weights);

/*
 * Run up to max_steps greedy steps, continuing from the current state.  For every emitted report row i:
 * idx_out[i] = sample index (first index among equal scores, np.argmax, utmos/select.py:48),
 * new_out[i] = new_count (utmos/select.py:49), score_out[i] = winning score after mask and weights
 * (may be NULL).  *n_out rows were written (<= max_steps); *stop_out is a UTMOS_STOP_* value.
 */
int utmos_select_steps(utmos_ctx *ctx, int64_t max_steps, int64_t *idx_out, int64_t *new_out,
                       double *score_out, int64_t *n_out, int *stop_out);

/*
 * Resume (no counterpart in the reference, which cannot resume a selection: SURVEY.md section 8, "next" item 4).
 * utmos_select_export copies out the state of the selection in progress: the working sample mask (picked samples are
 * 0), the live-row mask (utmos_info()[8] words; bit r = row r is not covered yet) and the report rows so far.
 * utmos_select_import starts a selection from such a state on a context that holds the same matrix: it recomputes
 * every gain from the matrix and the live mask, so utmos_select_steps continues with exactly the rows an uninterrupted
 * run would have produced.  Single-GPU contexts only.
 */
int utmos_select_export(utmos_ctx *ctx, uint8_t *mask_out, uint32_t *live_out, int64_t live_words, int64_t *idx_out,
                        int64_t *new_out, double *score_out, int64_t rows_cap, int64_t *n_rows_out, int64_t *tot_out,
                        int *stop_out);
int utmos_select_import(utmos_ctx *ctx, const uint8_t *mask, const double *weights, const uint32_t *live,
                        int64_t live_words, const int64_t *idx, const int64_t *new_count, const double *score,
                        int64_t n_rows, int64_t tot, int stop);

/*
 * utmos/convert.py:57-87 on a host int8 tensor gt[V][S][ploidy] (missing allele = -1).
 * packed_out: V x ceil(S/8) bytes, MSB-first (np.packbits); af_out[V] (NaN when no allele is called);
 * singleton_out[V] (may be NULL) = count(allele 1)==1 || count(allele 0)==1 (convert.py:58-60).
 */
int utmos_convert_gt(int device, const int8_t *gt, int64_t n_vars, int64_t n_samples, int64_t ploidy,
                     uint8_t *packed_out, double *af_out, int64_t *num_het_out, int64_t *num_hom_out,
                     uint8_t *singleton_out);

/* The same with flags.  UTMOS_CVT_DROP_SINGLETONS (`utmos convert --no-singleton`, utmos/convert.py:58-62): rows with
 * singleton_out[v] == 1 are left out of num_het / num_hom, as the reference drops them before it computes presence and
 * stats; their packed row and AF are still written (the caller compacts the rows by the flags).  One kernel launch per
 * block either way. */
#define UTMOS_CVT_DROP_SINGLETONS 1
int utmos_convert_gt_ex(int device, const int8_t *gt, int64_t n_vars, int64_t n_samples, int64_t ploidy,
                        uint8_t *packed_out, double *af_out, int64_t *num_het_out, int64_t *num_hom_out,
                        uint8_t *singleton_out, int flags);

/* CUDA-event milliseconds the K1 kernel launches of the last utmos_convert_gt call took (copies excluded). */
int utmos_convert_kernel_ms(double *ms_out);

/* ---- multi-GPU: rows sharded over ranks (one process per GPU), gains replicated ----
 * Every rank ingests ITS rows into its own context, then:
 *   utmos_rows            informative rows this rank kept (host all-reduces them -> UTMOS_OPT_GLOBAL_ROWS)
 *   utmos_finalize        local var_count / gains
 *   utmos_mgpu_layout     merged row numbering (see below)
 *   utmos_mgpu_export     allocates the exchange block (per-step delta inboxes + flags) and returns its 64-byte
 *                         CUDA IPC handle; the host all-gathers the handles (torch.distributed / MPI / files)
 *   utmos_mgpu_connect    maps every peer's block (NVLink P2P); handles = world x 64 bytes in rank order
 *   utmos_get_gains0 / utmos_set_gains0   step-0 gains out (local) and in (summed over ranks by the host)
 * After that utmos_select_begin / utmos_select_steps run the multi-GPU kernel: per step every rank finds the same
 * winner, retires its own rows into a delta vector, stores the delta straight into the peers' inboxes and
 * applies the sum.  All ranks must call utmos_select_steps with the same arguments; all return the same rows. */
int utmos_rows(utmos_ctx *ctx, int64_t *rows_out);
/* Merged row numbering for the hand-over to the replicated tail: once the part of the matrix that still scores is
 * sparse, every rank writes its live rows as edge-list entries into the exchange blocks of ALL ranks (NVLink
 * stores) and all ranks run the same single-CTA tail kernel.  row_base = sum over lower ranks of their rows
 * rounded up to 32, merged_rows = that sum over all ranks; allow_tail = 0 keeps the per-step exchange for the
 * whole selection (must be the same on every rank).  Call after utmos_set_gains0, before utmos_mgpu_export. */
int utmos_mgpu_layout(utmos_ctx *ctx, int64_t row_base, int64_t merged_rows, int allow_tail);
int utmos_mgpu_export(utmos_ctx *ctx, int rank, int world, uint8_t *handle_out /* 64 bytes */);
int utmos_mgpu_connect(utmos_ctx *ctx, const uint8_t *handles /* world x 64 bytes */);
int utmos_get_gains0(utmos_ctx *ctx, uint32_t *cnt_out, uint64_t *lo_out, uint64_t *hi_out);
int utmos_set_gains0(utmos_ctx *ctx, const uint32_t *cnt, const uint64_t *lo, const uint64_t *hi, int64_t global_rows);

/* ---- introspection (tests, benchmark) ---- */

/* Current per-sample state: gain counts (new_count each sample would get now) and unweighted, unmasked scores. */
int utmos_debug_gains(utmos_ctx *ctx, int64_t *count_out, double *score_out);

/* %globaltimer nanoseconds at which steps first..first+n-1 of the current selection were picked. */
int utmos_debug_step_times(utmos_ctx *ctx, int64_t first, int64_t n, int64_t *ns_out);

/* Profiling counters of the selection kernel: out16[0..3] = clock cycles spent (CTA 0, thread 0) in argmax,
 * barrier 1, cover, barrier 2, summed over steps; out16[4] = kernel launches; out16[11] = launches of the entry-divided
 * tail; out16[12] = steps of the per-step kernels that the reference-tie replay decided. */
int utmos_debug_counters(utmos_ctx *ctx, int64_t *out16);

/* Tunables.  UTMOS_OPT_REGAIN_ROWS: a pick that newly covers >= value rows triggers one streaming recompute of
 * all gains from the sample-major copy instead of per-bit subtraction (0 = never, -1 = default max(4096, V/170)). */
#define UTMOS_OPT_REGAIN_ROWS 1
#define UTMOS_OPT_GLOBAL_ROWS 4       /* multi-GPU: informative rows summed over all ranks; set before utmos_finalize */
#define UTMOS_OPT_STEP_TIMES 2        /* 1: record %globaltimer per pick for utmos_debug_step_times (adds latency) */
#define UTMOS_OPT_TAIL_ROWS 3         /* hand over to the list-driven tail kernel once a pick covers fewer rows (default 2048) */
#define UTMOS_OPT_LIST_BUDGET 11      /* before finalize: edge-list entries (live set bits) the list-driven tail may be built from
                                       * (0 = default 2^26, 2^28 for cohorts of more than 65,535 samples; 16 B each, 32 B with AF) */
#define UTMOS_OPT_TIE_ROW_CAP 12      /* UTMOS_F_REF_TIES: uncovered rows of a near-tie candidate the tail kernel replays itself (a power of two,
                                       * default what fits, at most 8,192); a candidate with more is handed to the per-step kernels */
#define UTMOS_OPT_TAIL_HEAVY_ROWS 10  /* list-driven tail: picks covering >= value rows run on the entry-divided 16-CTA cluster kernel
                                       * (gains sliced over the cluster's shared memories); lighter picks from one SM's shared memory (0 = never; -1 = default: 768 in count mode, 1 with AF or when the per-sample state does not fit one SM) */
#define UTMOS_OPT_TAIL_SINGLE_ROWS 5  /* tail kernel: 8-CTA owner-computes cluster until a pick covers fewer rows
                                         (default 0 = single CTA only) */
int utmos_set_option(utmos_ctx *ctx, int option, int64_t value);

/* info[0]=num_vars, [1]=row pitch bytes, [2]=has sample-major copy, [3]=device bytes in use,
 * [4]=fixed-point scale (AF flavours), [5]=AF values not exactly representable (count), [6]=kernels launched,
 * [7]=selection kernel flavour used: 0 step kernels, 1 grid-wide persistent, 2 one-cluster DSMEM,
 *   3 single-CTA list-driven tail (after a head run by flavour 1 or 2), 4 multi-GPU kernel (per-step exchange),
 *   5 multi-GPU head followed by the replicated tail;
 * [8]=32-bit words of the live-row mask (utmos_select_export / _import);
 * [9]=1 when the reference tie order (UTMOS_F_REF_TIES) is in effect (after utmos_finalize: it falls back to the exact
 *   order when the sample-major copy does not exist) */
int utmos_info(utmos_ctx *ctx, int64_t *info, int n);

/* Device-side (CUDA event) milliseconds accumulated since utmos_create / the last reset:
 * ms[0]=h2d copies, [1]=ingest kernels, [2]=transpose, [3]=column reduce / gain init, [4]=select loop;
 * parts of the select loop (host clock between stream synchronisations; the list build by its own CUDA events):
 * [5]=head launches, [6]=hand-over (edge-list build; multi-GPU: merge over NVLink), [7]=list-driven tail */
int utmos_timings(utmos_ctx *ctx, double *ms, int n, int reset);

/* --lowmem NEW.hdf5 writer (utmos/select.py:198-231): chunk i of the 'data' dataset = rows [i*chunk_rows, +chunk_rows)
 * of the packed .jl rows as S bool bytes each (or S float32 GT*AF each when af != NULL), zero padded past n_rows,
 * unpacked and LZF-compressed by `threads` host threads into dst + i*chunk_nbytes; sizes_out[i] = stored bytes,
 * masks_out[i] = 1 when the chunk is stored raw because LZF did not shrink it (hdf5 filter_mask bit 0). */
int utmos_h5_encode_chunks(const uint8_t *rows, int64_t n_rows, int64_t pitch, int64_t n_samples, const double *af,
                           int64_t chunk_rows, uint8_t *dst, int64_t *sizes_out, uint32_t *masks_out, int threads);

/* Device stopwatch (benchmark): start/stop synchronise the device and record a CUDA event each; *ms_out is the
 * event-to-event time, i.e. it covers every kernel and copy issued by any context on that device in between. */
int utmos_timer_start(int device);
int utmos_timer_stop(int device, double *ms_out);

/* ---- host-side codecs of the hdf5 chunk streamer (no device work) ----
 * hdf5 filter 32000 "lzf" as written by h5py for compression="lzf" (utmos/select.py:208-238).
 * decompress: returns decoded length or -1; compress: returns encoded length or 0 if it does not fit. */
int64_t utmos_lzf_decompress(const uint8_t *src, int64_t src_len, uint8_t *dst, int64_t dst_cap);
int64_t utmos_lzf_compress(const uint8_t *src, int64_t src_len, uint8_t *dst, int64_t dst_cap);

/* ---- host-side VCF text path of `utmos convert` (no device work; csrc/vcfio.cu) ----
 * Stands in for scikit-allel's parser, allel.read_vcf(fields=["calldata/GT","samples"]) (utmos/convert.py:50-53):
 * the int8 GT[V][S][2] tensor these produce is what utmos_convert_gt (K1) consumes.
 *   utmos_gz_size      uncompressed size of a gzip / BGZF buffer (exact for BGZF; ISIZE for one plain member), -1 if not gzip
 *   utmos_gz_inflate   inflate a whole buffer; BGZF blocks in parallel on `threads` host threads (0 = all cores)
 *   utmos_vcf_parse_gt tokenise up to max_variants data lines ('#' and empty lines skipped) into gt_out (missing = -1,
 *                      ploidy 2, phasing ignored, ploidy > 2 truncated, FORMAT without GT = all missing); *consumed_out
 *                      ends on a line boundary; a last line without newline is taken only when `final` is set */
int64_t utmos_gz_size(const uint8_t *src, int64_t len, int *is_bgzf_out);
int utmos_gz_inflate(const uint8_t *src, int64_t len, uint8_t *dst, int64_t dst_cap, int64_t *dst_len_out, int threads);
int utmos_vcf_parse_gt(const char *text, int64_t len, int64_t n_samples, int8_t *gt_out, int64_t max_variants,
                       int64_t *n_out, int64_t *consumed_out, int final, int threads);

/* ---- synthetic benchmark / test input generated in HBM (utmos_b200/synth.py holds the NumPy mirror) ----
 * Raw device buffers for the resident-input benchmark leg, and a deterministic cohort generator:
 * rows row0..row0+n_rows-1 in the .jl layout (MSB-first, pitch ceil(S/8)) + per-row AF, from host tables
 * cdf_thr[kmax+1] (allele-count CDF thresholds) and p_thr[kmax+1] (carrier probability thresholds). */
int utmos_device_alloc(int device, void **ptr_out, int64_t bytes);
int utmos_device_free(int device, void *ptr);
int utmos_device_to_host(int device, void *dst, const void *d_src, int64_t bytes);
int utmos_synth_packed_device(int device, uint64_t seed, int64_t row0, int64_t n_rows, int64_t n_samples,
                              const uint64_t *cdf_thr, const uint32_t *p_thr, int64_t kmax, uint8_t *d_rows,
                              double *d_af);

#ifdef __cplusplus
}
#endif
#endif /* UTMOS_B200_H */
