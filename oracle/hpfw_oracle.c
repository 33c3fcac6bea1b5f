/* oracle/hpfw_oracle.c
 *
 * TEST INFRASTRUCTURE ONLY. CPU restatement, in plain C, of the reference's hashprint feature-to-match path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this; the product
 * (hpfw_b200/, include/) never does.
 *
 * Parity status: stages a3..a7 and a12/a14 are PINNED — tests/test_oracle_cpu.py checks every function below against the
 * reference's own headers compiled into oracle/_ref/libhpfw_ref.so (see oracle/ref_build/) and against the committed
 * fixtures in tests/golden/ that were generated from that library. The CQT (a1/a2) lives in essentia, which is absent
 * from /root/reference: its restatement is oracle/nsgcq.py and is "parity unpinned" (see that file's header).
 *
 * Each function cites the reference file:line it restates (paths relative to /root/reference/include/hpfw/).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define ORC_BINS 121      /* spectrum/cqt.h:21  NumberBins                       */
#define ORC_CTX 20        /* live_song_id.h:16  FramesContext                    */
#define ORC_LAG 80        /* live_song_id.h:16  T                                */
#define ORC_NFILT 64      /* hashprint_handle.h:64  sizeof(uint64_t)*8           */
#define ORC_FRAME (ORC_BINS * ORC_CTX) /* hashprint_handle.h:60 FrameSize = 2420 */

/* ---- a3: spectrum/convert.h:7-25 -------------------------------------------------------------------------------
 * amplitude_to_db: p = x^2 ; power_to_db: mx = max(1e-10, max p); L = 10 log10(max(p,1e-10)) - 10 log10(mx);
 * out = max(L, max(L) - 80).  In/out column-major float[121 x cols]; computed in float like the reference (Real=float),
 * with the log evaluated in double and rounded (the reference's Eigen log10 is a float op under -ffast-math; the
 * difference is below 1e-5 dB and is covered by the stated tolerance in the tests). */
void orc_amplitude_to_db(float *s, int cols) {
    size_t n = (size_t)ORC_BINS * (size_t)cols;
    float mx = 1e-10f;
    for (size_t i = 0; i < n; ++i) {
        float p = s[i] * s[i];
        s[i] = p;
        if (p > mx) mx = p;
    }
    float lmx = 10.0f * (float)log10((double)mx);
    float top = -INFINITY;
    for (size_t i = 0; i < n; ++i) {
        float p = s[i] < 1e-10f ? 1e-10f : s[i];
        float l = 10.0f * (float)log10((double)p) - lmx;
        s[i] = l;
        if (l > top) top = l;
    }
    float floor_db = top - 80.0f;
    for (size_t i = 0; i < n; ++i)
        if (s[i] < floor_db) s[i] = floor_db;
}

/* ---- a4: core/hashprint_handle.h:79-93 ---------------------------------------------------------------------------
 * frames[(b*20 + c), t] = S[b, t + c], t = 0 .. cols-20. Row index is band-major with the context index inner: the
 * reference copies the 121x20 block into a RowMajor dynamic matrix and resize()s it to a column, which keeps the
 * row-major linear order (b*20+c). Output row-major float[2420 x (cols-19)] like Algo::Frames. Returns frame count. */
int orc_calc_frames(const float *s, int cols, float *frames) {
    int nf = cols - ORC_CTX + 1; /* cols - 2W + 1, W = 10 */
    if (nf <= 0) return 0;
    for (int b = 0; b < ORC_BINS; ++b)
        for (int c = 0; c < ORC_CTX; ++c) {
            float *row = frames + (size_t)(b * ORC_CTX + c) * (size_t)nf;
            for (int t = 0; t < nf; ++t) row[t] = s[(size_t)(t + c) * ORC_BINS + b];
        }
    return nf;
}

/* ---- a5: core/parallel_collector.h:57,127  y = filters * frames ----------------------------------------------------
 * filters column-major float[64 x 2420]; y column-major [64 x nf]. Frames are never materialised here (same numbers).
 * Two accumulations: float (sequential, like a scalar sgemm) and double (the rounding-free yardstick). */
int orc_project_f32(const float *s, int cols, const float *filters, float *y) {
    int nf = cols - ORC_CTX + 1;
    if (nf <= 0) return 0;
    for (int t = 0; t < nf; ++t)
        for (int f = 0; f < ORC_NFILT; ++f) {
            float acc = 0.0f;
            for (int b = 0; b < ORC_BINS; ++b)
                for (int c = 0; c < ORC_CTX; ++c)
                    acc += filters[(size_t)(b * ORC_CTX + c) * ORC_NFILT + f] * s[(size_t)(t + c) * ORC_BINS + b];
            y[(size_t)t * ORC_NFILT + f] = acc;
        }
    return nf;
}

int orc_project_f64(const float *s, int cols, const float *filters, double *y) {
    int nf = cols - ORC_CTX + 1;
    if (nf <= 0) return 0;
    for (int t = 0; t < nf; ++t)
        for (int f = 0; f < ORC_NFILT; ++f) {
            double acc = 0.0;
            for (int b = 0; b < ORC_BINS; ++b)
                for (int c = 0; c < ORC_CTX; ++c)
                    acc += (double)filters[(size_t)(b * ORC_CTX + c) * ORC_NFILT + f] *
                           (double)s[(size_t)(t + c) * ORC_BINS + b];
            y[(size_t)t * ORC_NFILT + f] = acc;
        }
    return nf;
}

/* ---- a6 + a7: core/hashprint_handle.h:115-142 ------------------------------------------------------------------------
 * bit(f,t) = (y[f,t] - y[f,t+80]) >= 0 ; hp[t] = sum_f bit(f,t) * 2^(63-f)  (filter 0 is the MSB for N = uint64_t:
 * c.reverse() visits filter 63 first with p = 0). Returns the word count ycols - 80 (0 if not positive). */
int orc_fingerprint_pack_f32(const float *y, int ycols, uint64_t *hp) {
    int n = ycols - ORC_LAG;
    if (n <= 0) return 0;
    for (int t = 0; t < n; ++t) {
        uint64_t w = 0;
        for (int f = 0; f < ORC_NFILT; ++f) {
            float d = y[(size_t)t * ORC_NFILT + f] - y[(size_t)(t + ORC_LAG) * ORC_NFILT + f];
            if (d >= 0.0f) w |= (uint64_t)1 << (63 - f);
        }
        hp[t] = w;
    }
    return n;
}

int orc_fingerprint_pack_f64(const double *y, int ycols, uint64_t *hp) {
    int n = ycols - ORC_LAG;
    if (n <= 0) return 0;
    for (int t = 0; t < n; ++t) {
        uint64_t w = 0;
        for (int f = 0; f < ORC_NFILT; ++f) {
            double d = y[(size_t)t * ORC_NFILT + f] - y[(size_t)(t + ORC_LAG) * ORC_NFILT + f];
            if (d >= 0.0) w |= (uint64_t)1 << (63 - f);
        }
        hp[t] = w;
    }
    return n;
}

/* |y[f,t] - y[f,t+80]| margins in double, for "which bits are allowed to differ" analysis in the tests. */
int orc_delta_f64(const double *y, int ycols, double *delta) {
    int n = ycols - ORC_LAG;
    if (n <= 0) return 0;
    for (int t = 0; t < n; ++t)
        for (int f = 0; f < ORC_NFILT; ++f)
            delta[(size_t)t * ORC_NFILT + f] = y[(size_t)t * ORC_NFILT + f] - y[(size_t)(t + ORC_LAG) * ORC_NFILT + f];
    return n;
}

/* ---- a8 (minus the spectrogram): core/parallel_collector.h:54-59 ------------------------------------------------------- */
int orc_hashprint_from_spectrogram(const float *s, int cols, const float *filters, uint64_t *hp) {
    int nf = cols - ORC_CTX + 1;
    if (nf - ORC_LAG <= 0) return 0;
    float *y = (float *)malloc(sizeof(float) * ORC_NFILT * (size_t)nf);
    orc_project_f32(s, cols, filters, y);
    int n = orc_fingerprint_pack_f32(y, nf, hp);
    free(y);
    return n;
}

/* ---- a12: audioproblems/live-song-id/storage.h:27-64  MemoryStorage::find ---------------------------------------------
 * DB = concatenated words + offsets[R+1]. Strict '<' everywhere: lowest offset within a track, earliest track overall.
 * k = min(k_query, n_ref) (storage.h:34-38); an empty reference yields distance 0 at offset 0; an empty DB leaves
 * {track -1, cnt SIZE_MAX, offset 0}. */
static inline uint64_t orc_best_in_track(const uint64_t *q, size_t kq, const uint64_t *r, size_t n, int64_t *best_off) {
    uint64_t best = UINT64_MAX;
    int64_t off = 0;
    size_t k = kq;
    if (n < k) k = n;
    for (size_t i = 0; i < n - k + 1; ++i) {
        uint64_t cnt = 0;
        for (size_t j = 0; j < k; ++j) cnt += (uint64_t)__builtin_popcountll(q[j] ^ r[i + j]);
        if (cnt < best) {
            best = cnt;
            off = (int64_t)i;
        }
    }
    *best_off = off;
    return best;
}

int64_t orc_find(const uint64_t *words, const int64_t *offsets, int n_tracks, const uint64_t *q, int kq, uint64_t *cnt,
                 int64_t *offset) {
    int64_t best_track = -1;
    *cnt = UINT64_MAX;
    *offset = 0;
    for (int r = 0; r < n_tracks; ++r) {
        int64_t off;
        uint64_t d = orc_best_in_track(q, (size_t)kq, words + offsets[r], (size_t)(offsets[r + 1] - offsets[r]), &off);
        if (d < *cnt) {
            *cnt = d;
            *offset = off;
            best_track = r;
        }
    }
    return best_track;
}

/* ---- a14: examples/python/liveid.ipynb:98-116,909-927 -------------------------------------------------------------------
 * Per-track best distance (and its lowest offset) for every track. Top-k = the k smallest (distance, track) pairs;
 * the notebook sorts (distance, label) tuples — we break distance ties by DB index (documented in DESIGN.md).
 * The notebook's Cython loop does NOT truncate the query when the reference is shorter (range(n-k+1) is empty and
 * 4294967294 is returned); we follow the C++ MemoryStorage semantics (truncate) for those tracks. */
void orc_per_track_best(const uint64_t *words, const int64_t *offsets, int n_tracks, const uint64_t *q, int kq,
                        uint64_t *dist, int64_t *off) {
    for (int r = 0; r < n_tracks; ++r)
        dist[r] = orc_best_in_track(q, (size_t)kq, words + offsets[r], (size_t)(offsets[r + 1] - offsets[r]), &off[r]);
}

/* out arrays have topk entries; unfilled entries: track -1, dist UINT64_MAX */
void orc_find_topk(const uint64_t *words, const int64_t *offsets, int n_tracks, const uint64_t *q, int kq, int topk,
                   int64_t *track_out, uint64_t *dist_out, int64_t *off_out) {
    uint64_t *dist = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(n_tracks > 0 ? n_tracks : 1));
    int64_t *off = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_tracks > 0 ? n_tracks : 1));
    unsigned char *used = (unsigned char *)calloc((size_t)(n_tracks > 0 ? n_tracks : 1), 1);
    orc_per_track_best(words, offsets, n_tracks, q, kq, dist, off);
    for (int i = 0; i < topk; ++i) {
        int best = -1;
        for (int r = 0; r < n_tracks; ++r)
            if (!used[r] && (best < 0 || dist[r] < dist[best])) best = r;
        if (best < 0) {
            track_out[i] = -1;
            dist_out[i] = UINT64_MAX;
            off_out[i] = 0;
        } else {
            used[best] = 1;
            track_out[i] = best;
            dist_out[i] = dist[best];
            off_out[i] = off[best];
        }
    }
    free(dist);
    free(off);
    free(used);
}

/* Query-parallel batch of orc_find_topk over host threads (the notebook's multiprocessing.Pool, liveid.ipynb:930-931).
 * This is the "port" CPU baseline; the "reference" baseline is oracle/_ref's ref_storage_find_batch. */
typedef struct {
    const uint64_t *words;
    const int64_t *offsets;
    int n_tracks;
    const uint64_t *qwords;
    const int64_t *qoffsets;
    int n_queries;
    int topk;
    int64_t *track_out;
    uint64_t *dist_out;
    int64_t *off_out;
    int *next;
    pthread_mutex_t *mtx;
} orc_batch_t;

static void *orc_batch_worker(void *p) {
    orc_batch_t *b = (orc_batch_t *)p;
    for (;;) {
        pthread_mutex_lock(b->mtx);
        int i = (*b->next)++;
        pthread_mutex_unlock(b->mtx);
        if (i >= b->n_queries) break;
        orc_find_topk(b->words, b->offsets, b->n_tracks, b->qwords + b->qoffsets[i],
                      (int)(b->qoffsets[i + 1] - b->qoffsets[i]), b->topk, b->track_out + (size_t)i * b->topk,
                      b->dist_out + (size_t)i * b->topk, b->off_out + (size_t)i * b->topk);
    }
    return NULL;
}

void orc_find_topk_batch(const uint64_t *words, const int64_t *offsets, int n_tracks, const uint64_t *qwords,
                         const int64_t *qoffsets, int n_queries, int topk, int64_t *track_out, uint64_t *dist_out,
                         int64_t *off_out, int n_threads) {
    int next = 0;
    pthread_mutex_t mtx;
    pthread_mutex_init(&mtx, NULL);
    orc_batch_t b = {words, offsets, n_tracks, qwords, qoffsets, n_queries, topk, track_out, dist_out, off_out, &next, &mtx};
    if (n_threads < 1) n_threads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    for (int t = 1; t < n_threads; ++t) pthread_create(&th[t], NULL, orc_batch_worker, &b);
    orc_batch_worker(&b);
    for (int t = 1; t < n_threads; ++t) pthread_join(th[t], NULL);
    free(th);
    pthread_mutex_destroy(&mtx);
}
