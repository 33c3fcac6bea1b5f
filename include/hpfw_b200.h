/* include/hpfw_b200.h — C ABI of the B200-native hashprint feature-to-match path.
 *
 * This is the drop-in boundary: plain pointers and sizes, int status codes, no C++/torch types. Everything above it
 * (include/hpfw/ C++ headers with the reference's class names, the ctypes mirror in hpfw_b200/, bench.py) calls only
 * these symbols; everything below it is hand-written sm_100a CUDA in hpfw_b200/csrc/.
 *
 * Reference interfaces replaced (paths relative to /root/reference):
 *   hpfw_db_build / hpfw_db_find*            <- db::MemoryStorage::build / find
 *                                               include/hpfw/audioproblems/live-song-id/storage.h:21-25, 27-64
 *   hpfw_db_find_topk*                       <- notebook top-10 ranking, examples/python/liveid.ipynb:98-116, 909-927
 *   hpfw_set_filters / hpfw_get_filters      <- ParallelCollector::filters (load/save),
 *                                               include/hpfw/core/parallel_collector.h:61-73, 78
 *   hpfw_hashprint_from_spectrogram*         <- HashprintHandle::calc_frames, `filters * frames`, calc_fingerprint,
 *                                               fingerprint_to_hashprint
 *                                               include/hpfw/core/hashprint_handle.h:79-93, 115-142;
 *                                               include/hpfw/core/parallel_collector.h:56-58, 126-128
 *   hpfw_cqt_spectrogram*                    <- spectrum::CQT::spectrogram after decoding (NSGConstantQ + abs/decimate +
 *                                               amplitude_to_db), include/hpfw/spectrum/cqt.h:54-84,
 *                                               include/hpfw/spectrum/convert.h:7-25
 *   hpfw_calc_hashprint_audio*               <- ParallelCollector::calc_hashprint on a decoded buffer,
 *                                               include/hpfw/core/parallel_collector.h:54-59
 *   par_collector_* / *_result_free          <- the reference's own C ABI, modules/python/parallel_collector_wrapper.hpp:12-38
 *                                               (declared in include/hpfw_b200_pyhpfw.h)
 *
 * Conventions
 *   - every function returns HPFW_OK (0) or a negative error code; hpfw_last_error() gives the thread-local message.
 *   - "host" entry points take host pointers and do their own H2D/D2H copies (what the reference-facing classes call);
 *     "_device" entry points take device pointers (payload already in HBM) plus host metadata, enqueue on `stream`
 *     (a cudaStream_t cast to void*; NULL = the context's own non-blocking stream; name the CUDA legacy default stream
 *     as cudaStreamLegacy, (void*)1) and do not synchronise.
 *   - there is NO CPU fallback: without a CUDA device every entry point fails with HPFW_ERR_CUDA.
 *   - layouts follow the reference: spectrogram = column-major float[121 x cols] (time-major in memory),
 *     filters = column-major float[64 x 2420] with row index band*20+context, hashprint word bit (63-f) = filter f.
 */
#ifndef HPFW_B200_H
#define HPFW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HPFW_OK 0
#define HPFW_ERR_CUDA (-1)      /* CUDA runtime / no device */
#define HPFW_ERR_ARG (-2)       /* bad argument */
#define HPFW_ERR_LIMIT (-3)     /* input exceeds a documented limit (key field widths, shared memory) */
#define HPFW_ERR_STATE (-4)     /* e.g. filters not set */
#define HPFW_ERR_SHORT (-5)     /* audio / spectrogram too short to yield a hashprint (reference: size_t underflow) */

/* geometry of the default instantiation HashprintHandle<uint64_t, spectrum::CQT<44100,96,24,121,3>, 20, 80>
 * (include/hpfw/audioproblems/live-song-id/live_song_id.h:16) */
#define HPFW_BINS 121
#define HPFW_CONTEXT 20
#define HPFW_LAG 80
#define HPFW_NFILTERS 64
#define HPFW_FRAME_SIZE (HPFW_BINS * HPFW_CONTEXT)

/* Packed match key: (distance << 40) | (track << 20) | offset. Integer '<' on keys is exactly the reference's
 * strict-'<' ordering: smaller distance, then earlier DB index, then lower offset (storage.h:50-60). */
#define HPFW_KEY_OFFSET_BITS 20
#define HPFW_KEY_TRACK_BITS 20
#define HPFW_KEY_DIST_SHIFT 40
#define HPFW_KEY_NONE UINT64_MAX
#define HPFW_MAX_TRACK_WORDS ((1 << HPFW_KEY_OFFSET_BITS) - 1)
#define HPFW_MAX_TRACKS ((1 << HPFW_KEY_TRACK_BITS) - 1)
#define HPFW_MAX_QUERY_WORDS 4096

typedef struct hpfw_ctx hpfw_ctx; /* one per GPU */
typedef struct hpfw_db hpfw_db;   /* a (shard of a) hashprint database resident in HBM */

/* = db::MemoryStorage::SearchResult (storage.h:11-15) with the filename replaced by the DB index.
 * No match (empty DB): track = -1, cnt = SIZE_MAX, offset = 0, as the reference's initial value (storage.h:28). */
typedef struct hpfw_match {
    int64_t track;
    uint64_t cnt;
    int64_t offset;
} hpfw_match;

const char *hpfw_last_error(void);
const char *hpfw_version(void);

int hpfw_device_count(void);   /* visible CUDA devices (0 without a driver / GPU) */
int hpfw_ctx_create(int device, hpfw_ctx **out);
void hpfw_ctx_destroy(hpfw_ctx *ctx);
int hpfw_ctx_device(const hpfw_ctx *ctx);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
uint64_t hpfw_ctx_launch_count(const hpfw_ctx *ctx);
int hpfw_ctx_synchronize(hpfw_ctx *ctx);
/* Per-kernel device timing for the roofline report: when enabled, every launch of a product kernel is bracketed by CUDA
 * events on the stream it is launched on. hpfw_ctx_timing_read waits for the recorded events, returns the accumulated
 * device time and launch count of kernel class `kernel`, and clears the accumulators of that class when reset != 0. */
#define HPFW_K_MATCH 0    /* match_kernel (stage 4) */
#define HPFW_K_TOPK 1     /* topk_kernel / merge_kernel */
#define HPFW_K_PROJECT 2  /* project_kernel (stages 2-3) */
#define HPFW_K_CQT 3      /* all CQT kernels (stage 1) */
#define HPFW_K_OTHER 4
#define HPFW_K_MATCH_TC 5 /* match_tc_kernel (stage 4 on the tensor cores) */
#define HPFW_K_COUNT 6
int hpfw_ctx_timing_enable(hpfw_ctx *ctx, int on);
int hpfw_ctx_timing_read(hpfw_ctx *ctx, int kernel, double *total_ms, uint64_t *launches, int reset);

/* ------------------------------------------------------------------------------------------------ matcher (stage 4) */
/* words: concatenated hashprints of n_tracks tracks, offsets[n_tracks+1]; track_base = global index of the first
 * track of this shard (0 for an unsharded DB). Host buffers; copied to HBM. */
int hpfw_db_build(hpfw_ctx *ctx, const uint64_t *words, const int64_t *offsets, int n_tracks, int64_t track_base,
                  hpfw_db **out);
/* same, words already on the device (copied device-to-device into the DB's own layout); offsets on the host */
int hpfw_db_build_device(hpfw_ctx *ctx, const uint64_t *d_words, const int64_t *offsets, int n_tracks,
                         int64_t track_base, void *stream, hpfw_db **out);
/* same from scattered device segments: DB track r = d_words[src_offsets[r] .. + lengths[r]) (host metadata) */
int hpfw_db_build_gather_device(hpfw_ctx *ctx, const uint64_t *d_words, const int64_t *src_offsets, const int64_t *lengths,
                                int n_tracks, int64_t track_base, void *stream, hpfw_db **out);
void hpfw_db_destroy(hpfw_db *db);
int hpfw_db_tracks(const hpfw_db *db);
int64_t hpfw_db_words(const hpfw_db *db);

/* MemoryStorage::find: one query, top-1. Host buffers. */
int hpfw_db_find(hpfw_db *db, const uint64_t *q, int k, hpfw_match *out);
/* n_queries queries (concatenated words + qoffsets[n_queries+1]); out[n_queries * topk], ranked by key. Host buffers. */
int hpfw_db_find_topk(hpfw_db *db, const uint64_t *qwords, const int64_t *qoffsets, int n_queries, int topk,
                      hpfw_match *out);
/* Device path: d_qwords on the device, qoffsets on the host (metadata); d_keys_out[n_queries * topk] packed keys. */
int hpfw_db_match_device(hpfw_db *db, const uint64_t *d_qwords, const int64_t *qoffsets, int n_queries, int topk,
                         uint64_t *d_keys_out, void *stream);
/* Query words already on the device, records on the host (synchronises `stream`): MemoryStorage::find for a batch whose
 * hashprints were extracted on this GPU. */
int hpfw_db_find_topk_device(hpfw_db *db, const uint64_t *d_qwords, const int64_t *qoffsets, int n_queries, int topk,
                             hpfw_match *out, void *stream);
/* which kernel runs the cross-correlation:
 *   0 = XOR + POPC on the integer pipes (matcher.cu);
 *   1 = exact GEMM on the tensor cores over +1/-1 signed bytes, tcgen05.mma.kind::i8 with s32 accumulation (match_tc.cu);
 *   3 = the same over +1.0/-1.0 e2m1 nibbles, tcgen05.mma.kind::mxf4.block_scale with unit scales and f32 accumulation
 *       (sums of +-1 stay far below 2^24: exact), twice the word rate of 1;
 *   2 (default) = per batch the cheapest mix: the queries, sorted by length, are split by a small dynamic programme into
 *       tensor-core groups of up to 128 (fp4 operands; int8 if the environment variable HPFW_MATCH_TC_F4=0 was set when the
 *       context was created; a group costs its longest query whether it holds 1 or 128) and queries that stay on the
 *       integer pipes (about 1/8 of a group each): a batch of >= ~9 equal-length queries goes to the tensor cores, a
 *       single find() stays on XOR + POPC.
 * All four give bit-identical results. The environment variable HPFW_MATCH_IMPL sets the initial value of a new context.
 * The tensor-core kernels are only used after a known-answer self-test against the integer-pipe kernel has passed on this
 * context (dot products of +-2^18 at 4,096 words, a long exact run followed by noise, 7 query lengths); when it fails,
 * impl 1 / 3 return HPFW_ERR_STATE here and from the match calls, and impl 2 runs every query on the integer pipes. */
int hpfw_set_match_impl(hpfw_ctx *ctx, int impl);
/* runs (once per context and encoding; later calls return the recorded verdict) the self-test described above:
 * HPFW_OK, or HPFW_ERR_STATE when the tensor-core kernel's keys differ from the integer-pipe kernel's. */
int hpfw_match_tc_selftest(hpfw_ctx *ctx, int fp4);
/* host-only (needs no device): the routing hpfw_db_match_device would apply to one batch under `impl` (fp4 != 0: fp4 operand
 * encoding for impl 2). group_out[q] = index of the tensor-core group query q joins, or -1 for the integer-pipe kernel. */
int hpfw_match_route(const int64_t *qoffsets, int n_queries, int impl, int fp4, int32_t *group_out);
/* Multi-GPU merge after an all-gather: d_keys_in[n_ranks][n_queries][topk] -> d_keys_out[n_queries][topk]. */
int hpfw_topk_merge_device(hpfw_ctx *ctx, const uint64_t *d_keys_in, int n_ranks, int n_queries, int topk,
                           uint64_t *d_keys_out, void *stream);
/* Host helper: unpack keys into hpfw_match records. */
void hpfw_keys_decode(const uint64_t *keys, int n, hpfw_match *out);
/* Algorithmic word-ops (one XOR64 + popcount64 each) of matching these queries against this DB:
 * sum over queries and tracks of (n_r - k + 1) * k with k = min(k_q, n_r). */
double hpfw_db_word_ops(const hpfw_db *db, const int64_t *qoffsets, int n_queries);

/* -------------------------------------------------------------------------- projection / threshold / pack (stages 2,3) */
int hpfw_set_filters(hpfw_ctx *ctx, const float *filters_colmajor_64x2420);
int hpfw_get_filters(hpfw_ctx *ctx, float *filters_colmajor_64x2420);
/* words produced for a spectrogram of `cols` columns: cols - 99 (<= 0: too short) */
int hpfw_hashprint_words_for_cols(int cols);
/* host: spectrogram[121 x cols] col-major -> hp_out[cols-99]; *n_out = words written */
int hpfw_hashprint_from_spectrogram(hpfw_ctx *ctx, const float *spectrogram, int cols, uint64_t *hp_out, int *n_out);
/* device, batched: n spectrograms concatenated in d_spectrograms, col_offsets[n+1] (host, in columns);
 * d_hp_out receives track i's words at hp_offsets[i] = sum_{j<i} max(cols_j - 99, 0). */
int hpfw_hashprint_from_spectrogram_device(hpfw_ctx *ctx, const float *d_spectrograms, const int64_t *col_offsets,
                                           int n, uint64_t *d_hp_out, void *stream);
/* which kernel runs stages 2-3: 4 (default) = 3 with two 128-frame tiles per CTA sharing the filter stream and one MMA-issuing
 * thread per tile (5: four tiles); 1 = tcgen05/TMEM/TMA implicit GEMM, tf32 inputs, fp32 accumulate, one context
 * block per tile addressed with row-offset descriptors (project_tc.cu); 3 = the same with fp16 inputs (the 10 mantissa bits of
 * tf32; half the MMAs and operand bytes); 2 = as 1 but reloading the window per tap; 0 = fp32 CUDA-core FFMA kernel
 * (project.cu), kept as the measurement baseline and for bit-level comparisons. */
int hpfw_set_projection_impl(hpfw_ctx *ctx, int impl);
/* diagnostic: the projection y = filters * frames itself (column-major float[64 x (cols-19)]), host buffers */
int hpfw_project(hpfw_ctx *ctx, const float *spectrogram, int cols, float *y_out);

/* ------------------------------------------------------------------------------ filter learning (index time, row a10) */
/* HashprintHandle::calc_cov + the accumulate of ParallelCollector::preprocess (hashprint_handle.h:96-102,
 * parallel_collector.h:92-97): the context keeps accum_cov (2420 x 2420 float) in HBM; each call adds the sample
 * covariance of one spectrogram's context frames. */
int hpfw_cov_reset(hpfw_ctx *ctx);
int hpfw_cov_set(hpfw_ctx *ctx, const float *accum_2420x2420);     /* e.g. cache/accum_cov.cereal payload */
int hpfw_cov_get(hpfw_ctx *ctx, float *accum_2420x2420);
/* the same with device buffers (d_accum: 2420 x 2420 floats in HBM), for the multi-GPU index: the per-GPU accumulators are
 * summed by one all-reduce between these two calls (hpfw_b200/sharded.py) */
int hpfw_cov_get_device(hpfw_ctx *ctx, float *d_accum_out, void *stream);
int hpfw_cov_set_device(hpfw_ctx *ctx, const float *d_accum, void *stream);
int hpfw_cov_add_spectrogram(hpfw_ctx *ctx, const float *spectrogram, int cols);                       /* host buffer */
int hpfw_cov_add_spectrogram_device(hpfw_ctx *ctx, const float *d_spectrogram, int cols, void *stream);
/* HashprintHandle::calc_filters (hashprint_handle.h:105-112): the 64 eigenvectors of the largest eigenvalues as rows of
 * a column-major 64 x 2420 matrix, descending eigenvalue. cov = host 2420 x 2420 symmetric matrix, or NULL for the
 * context's accumulator (any positive scale gives the same filters). Sign convention: the largest-magnitude component of
 * every filter is positive. eigenvalues_out (64 floats) may be NULL. */
int hpfw_calc_filters(hpfw_ctx *ctx, const float *cov, float *filters_out, float *eigenvalues_out);

/* ------------------------------------------------------------------------------------------------- CQT (stage 1) */
/* The band window NSGConstantQ is asked for is "hann" (cqt.h:58). essentia is not available to check which formula that
 * name selects, so the convention is a switch (default 0; environment variable HPFW_CQT_WINDOW=periodic|symmetric):
 *   0 = periodic Hann centred on the band, 0.5 + 0.5 cos(2 pi k / Lg), k = -floor(Lg/2) .. ceil(Lg/2)-1 (NSG toolbox);
 *   1 = symmetric Hann over the Lg taps, 0.5 - 0.5 cos(2 pi n / (Lg - 1)), n = k + floor(Lg/2).
 * Both are tested against oracle/nsgcq.py; which one equals essentia's is a to-be-verified item. */
int hpfw_set_cqt_window(hpfw_ctx *ctx, int window);
/* spectrogram columns for an n_samples-long buffer: M/3 + 1 (cqt.h:73); 0 if the design is degenerate */
int hpfw_cqt_cols(int64_t n_samples);
/* host: mono float audio (already at the analysis rate) -> dB spectrogram[121 x cols] col-major */
int hpfw_cqt_spectrogram(hpfw_ctx *ctx, const float *audio, int64_t n_samples, float *spectrogram_out, int *cols_out);
/* device: same, buffers in HBM; d_spectrogram_out must hold 121 * hpfw_cqt_cols(n_samples) floats */
int hpfw_cqt_spectrogram_device(hpfw_ctx *ctx, const float *d_audio, int64_t n_samples, float *d_spectrogram_out,
                                void *stream);
/* diagnostic: linear magnitudes |c_j[3i]| before the dB step */
int hpfw_cqt_magnitude(hpfw_ctx *ctx, const float *audio, int64_t n_samples, float *mag_out, int *cols_out);
/* host-only: the band layout of the NSG design for an n_samples-long input (FFT-bin units): pos[121], lg[121], M.
 * Needs no device. */
int hpfw_cqt_design(int64_t n_samples, int *pos_out, int *lg_out, int *m_out);
/* diagnostic: complex FFT (inverse != 0: unnormalised inverse) of n interleaved (re,im) floats through the same
 * two-pass mixed-radix kernels the CQT uses; n must split into two {2,3,5,7}-smooth factors <= 8192. Host buffers. */
int hpfw_fft_c2c(hpfw_ctx *ctx, const float *in_interleaved, float *out_interleaved, int n, int inverse);

/* ------------------------------------------------------------------------------ whole query path (stages 1-3, a8) */
int hpfw_hashprint_words_for_samples(int64_t n_samples);
int hpfw_calc_hashprint_audio(hpfw_ctx *ctx, const float *audio, int64_t n_samples, uint64_t *hp_out, int *n_out);
int hpfw_calc_hashprint_audio_device(hpfw_ctx *ctx, const float *d_audio, int64_t n_samples, uint64_t *d_hp_out,
                                     void *stream);

/* batched extraction (ParallelCollector::collect_fingerprints over many tracks, parallel_collector.h:115-137): n tracks
 * concatenated in d_audio, sample_offsets[n+1] (host; every offset even so that each track is 8-byte aligned);
 * d_hp_out receives track i's words at sum_{j<i} hpfw_hashprint_words_for_samples(len_j). */
int hpfw_calc_hashprint_audio_batch_device(hpfw_ctx *ctx, const float *d_audio, const int64_t *sample_offsets, int n,
                                           uint64_t *d_hp_out, void *stream);

/* 16-bit PCM input (pcm.cu): the samples a WAV file or a decoder delivers before MonoLoader's conversion to float
 * (cqt.h:45-52). Half the bytes of the float buffer across PCIe; `sample / 32768` is done on the device and is exact, so the
 * hashprints equal those of the float entry points on the converted samples bit for bit. Same conventions as above. */
int hpfw_pcm16_to_float_device(hpfw_ctx *ctx, const int16_t *d_pcm, int64_t n_samples, float *d_audio_out, void *stream);
int hpfw_cqt_spectrogram_pcm16(hpfw_ctx *ctx, const int16_t *pcm, int64_t n_samples, float *spectrogram_out, int *cols_out);
int hpfw_calc_hashprint_pcm16(hpfw_ctx *ctx, const int16_t *pcm, int64_t n_samples, uint64_t *hp_out, int *n_out);
int hpfw_calc_hashprint_pcm16_device(hpfw_ctx *ctx, const int16_t *d_pcm, int64_t n_samples, uint64_t *d_hp_out,
                                     void *stream);
int hpfw_calc_hashprint_pcm16_batch_device(hpfw_ctx *ctx, const int16_t *d_pcm, const int64_t *sample_offsets, int n,
                                           uint64_t *d_hp_out, void *stream);

/* ------------------------------------------------------------------ extraction stream (index / search runtime, xstream.cu) */
/* The device-resident pipeline behind ParallelCollector::prepare (parallel_collector.h:48-52, 82-137: preprocess fan-out over
 * files, covariance accumulate, collect_fingerprints over the cached spectrograms) and LiveSongIdentification::search
 * (live_song_id.h:35-54). Decode threads write samples straight into pinned staging slots; a submit enqueues
 * H2D -> [int16 -> float] -> CQT on one of 4 lane streams (-> covariance accumulate on a serial side stream) and KEEPS the dB
 * spectrogram in HBM; after the filters are known hpfw_xs_hash_kept runs the projection/threshold/pack over all resident
 * spectrograms in batched launches; the hashprints stay in HBM and become a database (hpfw_xs_build_db) or a query batch
 * (hpfw_xs_match) without visiting the host.
 * Threading: hpfw_xs_acquire, hpfw_xs_release and hpfw_xs_fetch_spectrogram may be called from any thread; all other calls
 * follow the context's rule (one call at a time per context). */
typedef struct hpfw_xs hpfw_xs;
#define HPFW_XS_PCM16 1        /* the slot holds int16 samples (else float32) */
#define HPFW_XS_COV 2          /* also add the track's frame covariance to the context's accumulator (hpfw_cov_*) */
#define HPFW_XS_SPECTROGRAM 4  /* (set by hpfw_xs_submit_spectrogram) the slot holds a dB spectrogram, not audio */
/* pinned host memory for the cache-writer side (hpfw_xs_fetch_spectrogram destinations) */
int hpfw_host_alloc(size_t bytes, void **out);
void hpfw_host_free(void *p);
/* slots: pinned staging slots (>= 2 per decode thread is plenty); slot_bytes: initial size of each (slots grow on demand).
 * HBM budget for resident spectrograms: HPFW_XS_ARENA_BYTES (default 70 % of the free device memory), in chunks of
 * HPFW_XS_CHUNK_BYTES (default 1 GiB). */
int hpfw_xs_create(hpfw_ctx *ctx, int slots, size_t slot_bytes, hpfw_xs **out);
void hpfw_xs_destroy(hpfw_xs *xs);
/* blocks until a staging slot is free; *host_ptr_out = its pinned buffer of at least `bytes` bytes (any thread) */
int hpfw_xs_acquire(hpfw_xs *xs, size_t bytes, int *slot_out, void **host_ptr_out);
/* give an acquired slot back without submitting it (decode failed; any thread) */
int hpfw_xs_release(hpfw_xs *xs, int slot);
/* Enqueue the slot's n_samples (flags: HPFW_XS_PCM16, HPFW_XS_COV); *track_out = index of the track in the stream (submission
 * order). Returns without synchronising; the slot goes back to the ring when its upload has completed (also on failure).
 * HPFW_ERR_SHORT: fewer than 100 spectrogram columns; HPFW_ERR_LIMIT: HBM budget for resident spectrograms exhausted. */
int hpfw_xs_submit(hpfw_xs *xs, int slot, int64_t n_samples, int flags, int *track_out);
/* the same for an already computed dB spectrogram (a cache/spectros/<stem> file, cache.h:30-33): slot = float[121 x cols] */
int hpfw_xs_submit_spectrogram(hpfw_xs *xs, int slot, int cols, int flags, int *track_out);
int hpfw_xs_tracks(hpfw_xs *xs);
int hpfw_xs_track_info(hpfw_xs *xs, int track, int *cols_out, int *words_out, int *resident_out);
/* copy a resident spectrogram to the host (waits for that track only; any thread; host_out best pinned: hpfw_host_alloc) */
int hpfw_xs_fetch_spectrogram(hpfw_xs *xs, int track, float *host_out);
/* wait for everything submitted so far (lanes, covariance) */
int hpfw_xs_wait(hpfw_xs *xs);
/* hash every resident spectrogram that has not been hashed yet with the context's current filters (batched launches) */
int hpfw_xs_hash_kept(hpfw_xs *xs);
/* free the resident spectrograms (their hashprints, if hashed, stay); the caller has joined its hpfw_xs_fetch_spectrogram threads */
int hpfw_xs_drop_kept(hpfw_xs *xs);
/* forget all tracks and hashprints, keep the buffers */
int hpfw_xs_reset(hpfw_xs *xs);
/* the hashprint store: device pointer, and per track its word offset (-1: not hashed) and length */
int hpfw_xs_hashprints_device(hpfw_xs *xs, const uint64_t **d_words_out, int64_t *offsets_out, int64_t *lengths_out);
int hpfw_xs_hashprint_host(hpfw_xs *xs, int track, uint64_t *out);
int hpfw_xs_hashprints_host(hpfw_xs *xs, uint64_t *out, int64_t n_words);   /* the first n_words of the store */
/* database from hashed tracks, DB index i = stream track order[i]; device-to-device */
int hpfw_xs_build_db(hpfw_xs *xs, const int *order, int n, int64_t track_base, hpfw_db **out);
/* all tracks of the stream as queries (store order) against db: out[tracks * topk] */
int hpfw_xs_match(hpfw_xs *xs, hpfw_db *db, int topk, hpfw_match *out);

/* ------------------------------------------------------------------------- the database sharded over several GPUs (shard.cu) */
/* DB partitioned by track into contiguous ranges balanced by matcher work, queries replicated, per-GPU top-k keys exchanged by
 * ONE in-place ncclAllGather and merged by a kernel: bit-identical to the single-GPU result for any number of shards (keys order
 * like the reference's strict '<' scan, storage.h:50-60). NCCL is called from inside the library (libnccl.so.2 is dlopen'ed on
 * first use; a process that already carries an NCCL, e.g. PyTorch's, shares it).
 *   local mode: one process drives n devices (hpfw_shard_create_local) - db::ShardedMemoryStorage in the C++ API;
 *   rank mode:  one process per GPU (hpfw_shard_create_rank); rank 0 calls hpfw_shard_unique_id and the launcher carries the
 *               128 bytes to the other ranks. */
typedef struct hpfw_shard hpfw_shard;
/* host-only (no device needed): bounds_out[n_shards + 1], shard s = tracks [bounds[s], bounds[s+1]) */
int hpfw_shard_plan(const int64_t *track_words, int n_tracks, int n_shards, int query_words, int *bounds_out);
/* the same with the work shared in proportion to speed[s] (e.g. 1 / measured match time of rank s on an equal split): the GPUs
 * of a node do not run at the same clock under their power caps, and the all-gather waits for the slowest rank */
int hpfw_shard_plan_weighted(const int64_t *track_words, int n_tracks, int n_shards, int query_words, const double *speed,
                             int *bounds_out);
int hpfw_shard_nccl_version(void);                     /* e.g. 22809; 0 when NCCL cannot be loaded */
int hpfw_shard_unique_id(void *id128_out);
int hpfw_shard_create_rank(hpfw_ctx *ctx, int rank, int world, const void *id128, hpfw_shard **out);
/* devices = NULL: devices 0 .. n_devices-1; the shard owns one hpfw_ctx per device (hpfw_shard_ctx) */
int hpfw_shard_create_local(const int *devices, int n_devices, hpfw_shard **out);
void hpfw_shard_destroy(hpfw_shard *s);
int hpfw_shard_world(const hpfw_shard *s);
int hpfw_shard_rank(const hpfw_shard *s);
hpfw_ctx *hpfw_shard_ctx(hpfw_shard *s, int local_index);
hpfw_db *hpfw_shard_db(hpfw_shard *s, int local_index);
/* rank mode: THIS rank's shard; track_base = global index of its first track */
int hpfw_shard_build_rank(hpfw_shard *s, const uint64_t *words, const int64_t *offsets, int n_tracks, int64_t track_base);
int hpfw_shard_build_rank_device(hpfw_shard *s, const uint64_t *d_words, const int64_t *offsets, int n_tracks,
                                 int64_t track_base, void *stream);
/* local mode: the WHOLE database (host words, or hashprints resident on device 0 as scattered segments in DB order); planned,
 * split and placed on the devices here. query_words_hint balances the shards (0 = 385, six seconds). */
int hpfw_shard_build(hpfw_shard *s, const uint64_t *words, const int64_t *offsets, int n_tracks, int query_words_hint);
int hpfw_shard_build_device(hpfw_shard *s, const uint64_t *d_words, const int64_t *src_offsets, const int64_t *lengths,
                            int n_tracks, int query_words_hint);
/* rank mode, everything on the device and on `stream`: local match into this rank's slot of the gather buffer, in-place
 * all-gather, merge; d_keys_out[n_queries * topk] is identical on every rank. Collective: every rank must call it. */
int hpfw_shard_match_device(hpfw_shard *s, const uint64_t *d_qwords, const int64_t *qoffsets, int n_queries, int topk,
                            uint64_t *d_keys_out, void *stream);
/* local mode: host queries in (or queries resident on device 0, broadcast over NVLink), host records out */
int hpfw_shard_find_topk(hpfw_shard *s, const uint64_t *qwords, const int64_t *qoffsets, int n_queries, int topk,
                         hpfw_match *out);
int hpfw_shard_find_topk_device(hpfw_shard *s, const uint64_t *d_qwords_dev0, const int64_t *qoffsets, int n_queries, int topk,
                                hpfw_match *out);
/* rank mode, index side: sum the contexts' covariance accumulators in place (parallel_collector.h:94-97 across GPUs);
 * byte-wise all-gather / broadcast for the hashprint and filter exchange */
int hpfw_shard_allreduce_cov(hpfw_shard *s, void *stream);
int hpfw_shard_allgather_device(hpfw_shard *s, const void *d_send, void *d_recv, size_t bytes_per_rank, void *stream);
/* ranks contribute bytes_per_rank[r] bytes each; d_recv receives them back to back in rank order (d_send may be NULL when this
 * rank's bytes already sit at their place in d_recv) */
int hpfw_shard_allgatherv_device(hpfw_shard *s, const void *d_send, void *d_recv, const size_t *bytes_per_rank, void *stream);
int hpfw_shard_broadcast_device(hpfw_shard *s, void *d_buf, size_t bytes, int root, void *stream);

/* ------------------------------------------------------------------------------------------------ measurement aids */
/* Pipe microbenchmark used to pin the matcher's roofline denominator: runs register-only loops and reports
 * lane-instructions per clock per SM for (0) POPC alone, (1) LOP3 alone, (2) the matcher's XOR/POPC/IADD3 mix,
 * expressed as 64-bit word-ops per clock per SM for (2). out[3]; sm_clock_mhz_out = clock observed during the run. */
int hpfw_microbench_pipes(hpfw_ctx *ctx, double *out, double *sm_clock_mhz_out);

#ifdef __cplusplus
}
#endif
#endif /* HPFW_B200_H */
