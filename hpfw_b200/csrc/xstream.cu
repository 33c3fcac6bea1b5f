// hpfw_b200/csrc/xstream.cu — the extraction stream: the native runtime behind ParallelCollector::prepare() and
// LiveSongIdentification::search() (include/hpfw/core/parallel_collector.h, .../live_song_id.h).
//
// The reference fans out over files with a task pool and keeps every intermediate on the host and on disk
// (/root/reference/include/hpfw/core/parallel_collector.h:85-108 preprocess: spectrogram -> frames -> covariance -> mutex ->
// cache file; :119-134 collect_fingerprints: re-read every cached spectrogram -> hashprint). Here the same two phases run as
// one device-resident pipeline:
//
//   decode threads --> pinned staging ring --H2D--> [int16 -> float] --> CQT (lane streams, cqt.cu) --> dB spectrogram KEPT in HBM
//                                                                              |--> covariance accumulate (one cov stream, learn.cu)
//                                                                              '--> optional D2H for the cache-file writer threads
//   ... filters learned (hpfw_calc_filters) ...
//   kept spectrograms --> ONE batched projection/threshold/pack launch per arena chunk (project_tc.cu) --> hashprints in HBM
//                     --> database built device-to-device (hpfw_db_build_gather_device) or matched in place (queries)
//
// Nothing above visits the host between the audio upload and the match result. The spectrograms of a whole shard stay
// resident (7 MB per 3-minute track; 12,500 tracks = 88 GB of the 180 GB HBM3e) in 1 GiB arena chunks, bounded by
// HPFW_XS_ARENA_BYTES (default: 70 % of the free device memory at creation); when the bound is hit hpfw_xs_submit returns
// HPFW_ERR_LIMIT and the host layer spills through the cache files, as the reference always does.
//
// Threading: hpfw_xs_acquire / hpfw_xs_release / hpfw_xs_fetch_spectrogram may be called from any thread (decode and cache
// writer threads); every other function follows the context's rule — one call at a time per context.
#include "common.cuh"

#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

using namespace hpfw_b200;

namespace {

constexpr int XS_SLOT_FREE = 0, XS_SLOT_HELD = 1, XS_SLOT_INFLIGHT = 2;

struct XsSlot {
    void *host = nullptr;
    size_t cap = 0;
    int state = XS_SLOT_FREE;
    cudaEvent_t h2d_done = nullptr;
};

struct XsTrack {
    int cols = 0, words = 0;
    int chunk = -1;              // arena chunk holding the spectrogram (-1: dropped)
    size_t spec_off = 0;         // floats into the chunk
    int64_t hp_off = -1;         // words into the hashprint store (-1: not hashed yet)
    cudaEvent_t ready = nullptr; // recorded when the spectrogram is complete
};

struct XsChunk {
    float *ptr = nullptr;
    size_t cap = 0, used = 0;    // floats
    int n_tracks = 0;            // spectrograms that live in this chunk
};

}  // namespace

struct hpfw_xs {
    hpfw_ctx *ctx = nullptr;
    std::mutex m;                       // ring + track table
    std::condition_variable cv;
    std::mutex fetch_m;                 // serialises the fetch stream
    std::vector<XsSlot> slots;
    int nl = 4;                         // lanes in use
    int next_lane = 0;
    DeviceBuffer lane_in[HPFW_CTX_LANES], lane_audio[HPFW_CTX_LANES];
    cudaStream_t cov_stream[2] = {nullptr, nullptr}, fetch_stream = nullptr;
    cudaEvent_t cov_done[2] = {nullptr, nullptr};
    int next_cov = 0;
    bool forked = false;                // the lanes have work that ctx->stream has not joined yet
    std::vector<XsChunk> chunks;
    size_t arena_budget = 0, arena_bytes = 0, chunk_floats = 0;
    std::vector<XsTrack> tracks;
    std::vector<cudaEvent_t> event_pool;
    DeviceBuffer hp;                    // hashprint store (words of all hashed tracks, store order)
    int64_t hp_words = 0;
    int hashed_upto = 0;                // tracks [0, hashed_upto) are hashed or dropped
};

static int xs_event(hpfw_xs *xs, cudaEvent_t *e) {
    if (!xs->event_pool.empty()) {
        *e = xs->event_pool.back();
        xs->event_pool.pop_back();
        return HPFW_OK;
    }
    HPFW_CUDA_TRY(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    return HPFW_OK;
}

// the lanes, the cov stream and the fetch stream start after everything the context's stream holds so far
static int xs_fork(hpfw_xs *xs) {
    if (xs->forked) return HPFW_OK;
    hpfw_ctx *ctx = xs->ctx;
    ctx->order_on(ctx->stream);
    HPFW_CUDA_TRY(cudaEventRecord(ctx->lane_fork, ctx->stream));
    for (int l = 0; l < xs->nl; ++l) HPFW_CUDA_TRY(cudaStreamWaitEvent(ctx->lane_stream[l], ctx->lane_fork, 0));
    for (int c = 0; c < 2; ++c) HPFW_CUDA_TRY(cudaStreamWaitEvent(xs->cov_stream[c], ctx->lane_fork, 0));
    xs->forked = true;
    return HPFW_OK;
}

// ... and the context's stream continues after everything the lanes and the cov stream hold
static int xs_join(hpfw_xs *xs) {
    if (!xs->forked) return HPFW_OK;
    hpfw_ctx *ctx = xs->ctx;
    ctx->order_on(ctx->stream);
    for (int l = 0; l < xs->nl; ++l) {
        HPFW_CUDA_TRY(cudaEventRecord(ctx->lane_join[l], ctx->lane_stream[l]));
        HPFW_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->lane_join[l], 0));
    }
    for (int c = 0; c < 2; ++c) {
        HPFW_CUDA_TRY(cudaEventRecord(xs->cov_done[c], xs->cov_stream[c]));
        HPFW_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, xs->cov_done[c], 0));
    }
    HPFW_TRY(cov_fold_side(ctx, ctx->stream));      // the second covariance slot's accumulator joins the context's
    xs->forked = false;
    return HPFW_OK;
}

// room for one spectrogram of `floats` floats; HPFW_ERR_LIMIT when the arena budget is exhausted
static int xs_arena_alloc(hpfw_xs *xs, size_t floats, int *chunk_out, size_t *off_out) {
    if (!xs->chunks.empty()) {
        XsChunk &c = xs->chunks.back();
        if (c.used + floats <= c.cap) {
            *chunk_out = int(xs->chunks.size()) - 1;
            *off_out = c.used;
            c.used += floats;
            c.n_tracks++;
            return HPFW_OK;
        }
    }
    if (!xs->chunks.empty() && xs->chunks.back().n_tracks == 0) {
        // an empty chunk that is too small for this spectrogram (left over by hpfw_xs_drop_kept) makes room for a larger one
        cudaFree(xs->chunks.back().ptr);
        xs->arena_bytes -= xs->chunks.back().cap * sizeof(float);
        xs->chunks.pop_back();
    }
    const size_t want = std::max(xs->chunk_floats, floats);
    if (xs->arena_bytes + want * sizeof(float) > xs->arena_budget)
        HPFW_FAIL(HPFW_ERR_LIMIT, "extraction stream: the HBM budget for resident spectrograms is exhausted (%zu MiB in %zu "
                  "chunks, %zu tracks; HPFW_XS_ARENA_BYTES = %zu): spill through the cache files (hpfw_xs_drop_kept) or raise it",
                  xs->arena_bytes >> 20, xs->chunks.size(), xs->tracks.size(), xs->arena_budget);
    XsChunk c;
    HPFW_CUDA_TRY(cudaMalloc(&c.ptr, want * sizeof(float)));
    c.cap = want;
    c.used = floats;
    c.n_tracks = 1;
    xs->arena_bytes += want * sizeof(float);
    xs->chunks.push_back(c);
    *chunk_out = int(xs->chunks.size()) - 1;
    *off_out = 0;
    return HPFW_OK;
}

static void xs_arena_undo(hpfw_xs *xs, int chunk, size_t floats) {
    XsChunk &c = xs->chunks[size_t(chunk)];
    c.used -= floats;
    c.n_tracks--;
}

extern "C" {

int hpfw_host_alloc(size_t bytes, void **out) {
    if (!out) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_host_alloc: out is NULL");
    *out = nullptr;
    HPFW_CUDA_TRY(cudaMallocHost(out, std::max<size_t>(bytes, 1)));
    return HPFW_OK;
}

void hpfw_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

int hpfw_xs_create(hpfw_ctx *ctx, int slots, size_t slot_bytes, hpfw_xs **out) {
    if (!ctx || !out || slots < 1 || slots > 4096) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_xs_create: bad argument");
    *out = nullptr;
    DeviceGuard g(ctx->device);
    HPFW_TRY(ctx_lanes_init(ctx));
    hpfw_xs *xs = new hpfw_xs();
    xs->ctx = ctx;
    xs->slots.resize(size_t(slots));
    int st = HPFW_OK;
    for (auto &s : xs->slots) {
        if (cudaEventCreateWithFlags(&s.h2d_done, cudaEventDisableTiming) != cudaSuccess) st = HPFW_ERR_CUDA;
        if (st == HPFW_OK && slot_bytes) {
            if (cudaMallocHost(&s.host, slot_bytes) != cudaSuccess) st = HPFW_ERR_CUDA;
            else s.cap = slot_bytes;
        }
    }
    for (int c = 0; c < 2; ++c)
        if (st == HPFW_OK && cudaStreamCreateWithFlags(&xs->cov_stream[c], cudaStreamNonBlocking) != cudaSuccess) st = HPFW_ERR_CUDA;
    if (st == HPFW_OK && cudaStreamCreateWithFlags(&xs->fetch_stream, cudaStreamNonBlocking) != cudaSuccess) st = HPFW_ERR_CUDA;
    for (int c = 0; c < 2; ++c)
        if (st == HPFW_OK && cudaEventCreateWithFlags(&xs->cov_done[c], cudaEventDisableTiming) != cudaSuccess) st = HPFW_ERR_CUDA;
    size_t free_b = 0, total_b = 0;
    if (st == HPFW_OK && cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) st = HPFW_ERR_CUDA;
    if (st != HPFW_OK) {
        set_error("hpfw_xs_create: %s", cudaGetErrorString(cudaGetLastError()));
        hpfw_xs_destroy(xs);
        return st;
    }
    xs->arena_budget = size_t(double(free_b) * 0.7);
    if (const char *env = getenv("HPFW_XS_ARENA_BYTES")) xs->arena_budget = strtoull(env, nullptr, 10);
    size_t chunk_bytes = size_t(1) << 30;
    if (const char *env = getenv("HPFW_XS_CHUNK_BYTES")) chunk_bytes = std::max<size_t>(1 << 16, strtoull(env, nullptr, 10));
    xs->chunk_floats = std::min(chunk_bytes, std::max<size_t>(xs->arena_budget, 1 << 16)) / sizeof(float);
    xs->nl = 4;
    if (const char *env = getenv("HPFW_CQT_LANES")) xs->nl = std::max(1, std::min(HPFW_CTX_LANES, atoi(env)));
    *out = xs;
    return HPFW_OK;
}

void hpfw_xs_destroy(hpfw_xs *xs) {
    if (!xs) return;
    DeviceGuard g(xs->ctx->device);
    cudaDeviceSynchronize();
    for (auto &s : xs->slots) {
        if (s.host) cudaFreeHost(s.host);
        if (s.h2d_done) cudaEventDestroy(s.h2d_done);
    }
    for (auto &c : xs->chunks)
        if (c.ptr) cudaFree(c.ptr);
    for (auto &t : xs->tracks)
        if (t.ready) cudaEventDestroy(t.ready);
    for (auto e : xs->event_pool) cudaEventDestroy(e);
    for (int l = 0; l < HPFW_CTX_LANES; ++l) {
        xs->lane_in[l].release();
        xs->lane_audio[l].release();
    }
    xs->hp.release();
    for (int c = 0; c < 2; ++c) {
        if (xs->cov_stream[c]) cudaStreamDestroy(xs->cov_stream[c]);
        if (xs->cov_done[c]) cudaEventDestroy(xs->cov_done[c]);
    }
    if (xs->fetch_stream) cudaStreamDestroy(xs->fetch_stream);
    delete xs;
}

int hpfw_xs_acquire(hpfw_xs *xs, size_t bytes, int *slot_out, void **host_ptr_out) {
    if (!xs || !slot_out || !host_ptr_out) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_xs_acquire: NULL argument");
    DeviceGuard g(xs->ctx->device);
    std::unique_lock<std::mutex> lk(xs->m);
    int pick = -1;
    for (;;) {
        for (size_t i = 0; i < xs->slots.size() && pick < 0; ++i)
            if (xs->slots[i].state == XS_SLOT_FREE) pick = int(i);
        if (pick < 0) {
            // a slot whose upload has completed is free again
            int oldest = -1;
            for (size_t i = 0; i < xs->slots.size() && pick < 0; ++i) {
                if (xs->slots[i].state != XS_SLOT_INFLIGHT) continue;
                if (cudaEventQuery(xs->slots[i].h2d_done) == cudaSuccess) pick = int(i);
                else if (oldest < 0) oldest = int(i);
            }
            cudaGetLastError();      // cudaErrorNotReady is not an error
            if (pick < 0 && oldest >= 0) {
                cudaEvent_t e = xs->slots[size_t(oldest)].h2d_done;
                lk.unlock();
                cudaEventSynchronize(e);
                lk.lock();
                continue;            // states may have changed meanwhile: look again
            }
        }
        if (pick >= 0) break;
        xs->cv.wait(lk);             // every slot is held by a decoder: wait for a submit / release
    }
    XsSlot &s = xs->slots[size_t(pick)];
    s.state = XS_SLOT_HELD;
    lk.unlock();
    if (s.cap < bytes) {             // grow outside the lock: only this thread owns the slot now
        if (s.host) cudaFreeHost(s.host);
        s.host = nullptr;
        s.cap = 0;
        const size_t want = bytes + bytes / 4 + 4096;
        if (cudaMallocHost(&s.host, want) != cudaSuccess) {
            lk.lock();
            s.state = XS_SLOT_FREE;
            xs->cv.notify_all();
            HPFW_FAIL(HPFW_ERR_CUDA, "hpfw_xs_acquire: cannot pin %zu bytes of host memory", want);
        }
        s.cap = want;
    }
    *slot_out = pick;
    *host_ptr_out = s.host;
    return HPFW_OK;
}

int hpfw_xs_release(hpfw_xs *xs, int slot) {
    if (!xs || slot < 0 || slot >= int(xs->slots.size())) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_xs_release: bad slot");
    std::lock_guard<std::mutex> lk(xs->m);
    xs->slots[size_t(slot)].state = XS_SLOT_FREE;
    xs->cv.notify_all();
    return HPFW_OK;
}

static int xs_submit_common(hpfw_xs *xs, int slot, int64_t n_samples, int cols, int flags, int *track_out) {
    hpfw_ctx *ctx = xs->ctx;
    const bool is_spec = (flags & HPFW_XS_SPECTROGRAM) != 0;
    const int words = hpfw_hashprint_words_for_cols(cols);
    XsSlot &sl = xs->slots[size_t(slot)];
    auto fail_slot = [&]() {
        std::lock_guard<std::mutex> lk(xs->m);
        sl.state = XS_SLOT_FREE;
        xs->cv.notify_all();
    };
    if (words <= 0) {
        fail_slot();
        HPFW_FAIL(HPFW_ERR_SHORT, "%lld %s give %d spectrogram columns; at least 100 are needed for one hashprint word",
                  (long long)(is_spec ? cols : n_samples), is_spec ? "columns" : "samples", cols);
    }
    const int track = int(xs->tracks.size());
    const size_t floats = size_t(cols) * HPFW_BINS;
    int chunk = -1;
    size_t off = 0;
    int st;
    {
        std::lock_guard<std::mutex> lk(xs->m);       // hpfw_xs_fetch_spectrogram reads the chunk table from other threads
        st = xs_arena_alloc(xs, floats, &chunk, &off);
    }
    if (st != HPFW_OK) {
        fail_slot();
        return st;
    }
    XsTrack t;
    t.cols = cols;
    t.words = words;
    t.chunk = chunk;
    t.spec_off = off;
    float *d_spec = xs->chunks[size_t(chunk)].ptr + off;
    auto bail = [&](int code) {
        {
            std::lock_guard<std::mutex> lk(xs->m);
            xs_arena_undo(xs, chunk, floats);
        }
        if (t.ready) xs->event_pool.push_back(t.ready);
        fail_slot();
        return code;
    };
    st = xs_event(xs, &t.ready);
    if (st == HPFW_OK) st = xs_fork(xs);
    if (st != HPFW_OK) return bail(st);
    const int lane = xs->next_lane++ % xs->nl;
    cudaStream_t s = ctx->lane_stream[lane];
    cudaError_t e = cudaSuccess;
    if (is_spec) {
        e = cudaMemcpyAsync(d_spec, sl.host, floats * sizeof(float), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaEventRecord(sl.h2d_done, s);
    } else {
        const bool pcm = (flags & HPFW_XS_PCM16) != 0;
        const size_t in_bytes = size_t(n_samples) * (pcm ? sizeof(int16_t) : sizeof(float));
        DeviceBuffer &din = xs->lane_in[lane];
        if (din.cap < in_bytes + 64) st = din.reserve(in_bytes + in_bytes / 2 + 64);
        if (st == HPFW_OK && pcm && xs->lane_audio[lane].cap < sizeof(float) * size_t(n_samples + 8))
            st = xs->lane_audio[lane].reserve(sizeof(float) * size_t(n_samples + 8) * 3 / 2);
        if (st != HPFW_OK) return bail(st);
        e = cudaMemcpyAsync(din.ptr, sl.host, in_bytes, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaEventRecord(sl.h2d_done, s);
        const float *d_audio = din.as<float>();
        if (e == cudaSuccess && pcm) {
            st = pcm16_convert(ctx, din.as<int16_t>(), xs->lane_audio[lane].as<float>(), n_samples, s);
            d_audio = xs->lane_audio[lane].as<float>();
        }
        if (e == cudaSuccess && st == HPFW_OK) st = cqt_run_lane(ctx, d_audio, n_samples, d_spec, s, lane);
    }
    if (e != cudaSuccess) {
        set_error("hpfw_xs_submit: %s", cudaGetErrorString(e));
        st = HPFW_ERR_CUDA;
    }
    if (st == HPFW_OK && cudaEventRecord(t.ready, s) != cudaSuccess) st = HPFW_ERR_CUDA;
    if (st == HPFW_OK && (flags & HPFW_XS_COV)) {
        // two covariance slots (own scratch and accumulator each, learn.cu) on two streams that trail the lanes: the nine
        // small kernels of one track overlap the next track's
        const int c = xs->next_cov++ & 1;
        if (cudaStreamWaitEvent(xs->cov_stream[c], t.ready, 0) != cudaSuccess) st = HPFW_ERR_CUDA;
        if (st == HPFW_OK) st = cov_add_device(ctx, d_spec, cols, xs->cov_stream[c], c);
    }
    if (st != HPFW_OK) return bail(st);
    {
        std::lock_guard<std::mutex> lk(xs->m);
        sl.state = XS_SLOT_INFLIGHT;
        xs->tracks.push_back(t);
        xs->cv.notify_all();
    }
    if (track_out) *track_out = track;
    return HPFW_OK;
}

int hpfw_xs_submit(hpfw_xs *xs, int slot, int64_t n_samples, int flags, int *track_out) {
    if (!xs || slot < 0 || slot >= int(xs->slots.size()) || n_samples < 0)
        HPFW_FAIL(HPFW_ERR_ARG, "hpfw_xs_submit: bad argument");
    DeviceGuard g(xs->ctx->device);
    if (xs->slots[size_t(slot)].state != XS_SLOT_HELD) HPFW_FAIL(HPFW_ERR_STATE, "hpfw_xs_submit: slot %d was not acquired", slot);
    return xs_submit_common(xs, slot, n_samples, hpfw_cqt_cols(n_samples), flags & ~HPFW_XS_SPECTROGRAM, track_out);
}

int hpfw_xs_submit_spectrogram(hpfw_xs *xs, int slot, int cols, int flags, int *track_out) {
    if (!xs || slot < 0 || slot >= int(xs->slots.size()) || cols < 0)
        HPFW_FAIL(HPFW_ERR_ARG, "hpfw_xs_submit_spectrogram: bad argument");
    DeviceGuard g(xs->ctx->device);
    if (xs->slots[size_t(slot)].state != XS_SLOT_HELD)
        HPFW_FAIL(HPFW_ERR_STATE, "hpfw_xs_submit_spectrogram: slot %d was not acquired", slot);
    return xs_submit_common(xs, slot, 0, cols, flags | HPFW_XS_SPECTROGRAM, track_out);
}

int hpfw_xs_tracks(hpfw_xs *xs) {
    if (!xs) return 0;
    std::lock_guard<std::mutex> lk(xs->m);
    return int(xs->tracks.size());
}

int hpfw_xs_track_info(hpfw_xs *xs, int track, int *cols_out, int *words_out, int *resident_out) {
    if (!xs) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_xs_track_info: xs is NULL");
    std::lock_guard<std::mutex> lk(xs->m);
    if (track < 0 || track >= int(xs->tracks.size())) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_xs_track_info: no track %d", track);
    const XsTrack &t = xs->tracks[size_t(track)];
    if (cols_out) *cols_out = t.cols;
    if (words_out) *words_out = t.words;
    if (resident_out) *resident_out = t.chunk >= 0;
    return HPFW_OK;
}

int hpfw_xs_fetch_spectrogram(hpfw_xs *xs, int track, float *host_out) {
    if (!xs || !host_out) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_xs_fetch_spectrogram: NULL argument");
    DeviceGuard g(xs->ctx->device);
    XsTrack t;
    const float *src = nullptr;
    {
        std::lock_guard<std::mutex> lk(xs->m);
        if (track < 0 || track >= int(xs->tracks.size())) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_xs_fetch_spectrogram: no track %d", track);
        t = xs->tracks[size_t(track)];
        if (t.chunk < 0) HPFW_FAIL(HPFW_ERR_STATE, "hpfw_xs_fetch_spectrogram: the spectrogram of track %d was dropped", track);
        src = xs->chunks[size_t(t.chunk)].ptr + t.spec_off;
    }
    std::lock_guard<std::mutex> fl(xs->fetch_m);
    HPFW_CUDA_TRY(cudaStreamWaitEvent(xs->fetch_stream, t.ready, 0));
    HPFW_CUDA_TRY(cudaMemcpyAsync(host_out, src, sizeof(float) * size_t(t.cols) * HPFW_BINS, cudaMemcpyDeviceToHost,
                                  xs->fetch_stream));
    HPFW_CUDA_TRY(cudaStreamSynchronize(xs->fetch_stream));
    return HPFW_OK;
}

int hpfw_xs_wait(hpfw_xs *xs) {
    if (!xs) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_xs_wait: xs is NULL");
    DeviceGuard g(xs->ctx->device);
    HPFW_TRY(xs_join(xs));
    HPFW_CUDA_TRY(cudaStreamSynchronize(xs->ctx->stream));
    return HPFW_OK;
}

int hpfw_xs_hash_kept(hpfw_xs *xs) {
    if (!xs) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_xs_hash_kept: xs is NULL");
    hpfw_ctx *ctx = xs->ctx;
    DeviceGuard g(ctx->device);
    if (!ctx->have_filters) HPFW_FAIL(HPFW_ERR_STATE, "filters not set: call hpfw_set_filters first");
    HPFW_TRY(xs_join(xs));
    const int n = int(xs->tracks.size());
    int64_t need = xs->hp_words;
    for (int t = xs->hashed_upto; t < n; ++t)
        if (xs->tracks[size_t(t)].chunk >= 0) need += xs->tracks[size_t(t)].words;
    if (size_t(need) * sizeof(uint64_t) > xs->hp.cap) {
        // grow the store, keeping the words hashed so far
        DeviceBuffer bigger;
        HPFW_TRY(bigger.reserve(size_t(need) * sizeof(uint64_t) * 3 / 2 + 4096));
        if (xs->hp_words)
            HPFW_CUDA_TRY(cudaMemcpyAsync(bigger.ptr, xs->hp.ptr, sizeof(uint64_t) * size_t(xs->hp_words),
                                          cudaMemcpyDeviceToDevice, ctx->stream));
        HPFW_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        xs->hp.release();
        xs->hp = bigger;
    }
    // runs of consecutive tracks that are contiguous in one chunk -> one batched projection launch each
    int t = xs->hashed_upto;
    while (t < n) {
        if (xs->tracks[size_t(t)].chunk < 0) {
            ++t;
            continue;
        }
        const int chunk = xs->tracks[size_t(t)].chunk;
        const size_t base = xs->tracks[size_t(t)].spec_off;
        std::vector<int64_t> co{0};
        int u = t;
        size_t expect = base;
        while (u < n && xs->tracks[size_t(u)].chunk == chunk && xs->tracks[size_t(u)].spec_off == expect) {
            co.push_back(co.back() + xs->tracks[size_t(u)].cols);
            expect += size_t(xs->tracks[size_t(u)].cols) * HPFW_BINS;
            ++u;
        }
        HPFW_TRY(hpfw_hashprint_from_spectrogram_device(ctx, xs->chunks[size_t(chunk)].ptr + base, co.data(), u - t,
                                                        xs->hp.as<uint64_t>() + xs->hp_words, ctx->stream));
        for (int v = t; v < u; ++v) {
            xs->tracks[size_t(v)].hp_off = xs->hp_words;
            xs->hp_words += xs->tracks[size_t(v)].words;
        }
        t = u;
    }
    xs->hashed_upto = n;
    return HPFW_OK;
}

int hpfw_xs_drop_kept(hpfw_xs *xs) {
    if (!xs) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_xs_drop_kept: xs is NULL");
    DeviceGuard g(xs->ctx->device);
    HPFW_TRY(xs_join(xs));
    HPFW_CUDA_TRY(cudaStreamSynchronize(xs->ctx->stream));
    {
        std::lock_guard<std::mutex> fl(xs->fetch_m);
        HPFW_CUDA_TRY(cudaStreamSynchronize(xs->fetch_stream));
    }
    std::lock_guard<std::mutex> lk(xs->m);
    for (auto &t : xs->tracks) t.chunk = -1;
    xs->hashed_upto = int(xs->tracks.size());
    // keep one chunk for the next batch, give the rest back
    for (size_t i = 1; i < xs->chunks.size(); ++i) {
        cudaFree(xs->chunks[i].ptr);
        xs->arena_bytes -= xs->chunks[i].cap * sizeof(float);
    }
    if (xs->chunks.size() > 1) xs->chunks.resize(1);
    for (auto &c : xs->chunks) {
        c.used = 0;
        c.n_tracks = 0;
    }
    return HPFW_OK;
}

int hpfw_xs_reset(hpfw_xs *xs) {
    if (!xs) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_xs_reset: xs is NULL");
    HPFW_TRY(hpfw_xs_drop_kept(xs));
    std::lock_guard<std::mutex> lk(xs->m);
    for (auto &t : xs->tracks)
        if (t.ready) xs->event_pool.push_back(t.ready);
    xs->tracks.clear();
    xs->hp_words = 0;
    xs->hashed_upto = 0;
    return HPFW_OK;
}

int hpfw_xs_hashprints_device(hpfw_xs *xs, const uint64_t **d_words_out, int64_t *offsets_out, int64_t *lengths_out) {
    if (!xs) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_xs_hashprints_device: xs is NULL");
    std::lock_guard<std::mutex> lk(xs->m);
    if (d_words_out) *d_words_out = xs->hp.as<uint64_t>();
    for (size_t t = 0; t < xs->tracks.size(); ++t) {
        if (offsets_out) offsets_out[t] = xs->tracks[t].hp_off;
        if (lengths_out) lengths_out[t] = xs->tracks[t].hp_off >= 0 ? xs->tracks[t].words : 0;
    }
    return HPFW_OK;
}

int hpfw_xs_hashprint_host(hpfw_xs *xs, int track, uint64_t *out) {
    if (!xs || !out) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_xs_hashprint_host: NULL argument");
    hpfw_ctx *ctx = xs->ctx;
    DeviceGuard g(ctx->device);
    if (track < 0 || track >= int(xs->tracks.size())) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_xs_hashprint_host: no track %d", track);
    const XsTrack &t = xs->tracks[size_t(track)];
    if (t.hp_off < 0) HPFW_FAIL(HPFW_ERR_STATE, "hpfw_xs_hashprint_host: track %d has not been hashed", track);
    ctx->order_on(ctx->stream);
    HPFW_CUDA_TRY(cudaMemcpyAsync(out, xs->hp.as<uint64_t>() + t.hp_off, sizeof(uint64_t) * size_t(t.words),
                                  cudaMemcpyDeviceToHost, ctx->stream));
    HPFW_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return HPFW_OK;
}

int hpfw_xs_hashprints_host(hpfw_xs *xs, uint64_t *out, int64_t n_words) {
    if (!xs || (!out && n_words > 0)) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_xs_hashprints_host: NULL argument");
    hpfw_ctx *ctx = xs->ctx;
    DeviceGuard g(ctx->device);
    if (n_words > xs->hp_words) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_xs_hashprints_host: the store holds %lld words", (long long)xs->hp_words);
    if (n_words <= 0) return HPFW_OK;
    ctx->order_on(ctx->stream);
    HPFW_CUDA_TRY(cudaMemcpyAsync(out, xs->hp.ptr, sizeof(uint64_t) * size_t(n_words), cudaMemcpyDeviceToHost, ctx->stream));
    HPFW_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return HPFW_OK;
}

int hpfw_xs_build_db(hpfw_xs *xs, const int *order, int n, int64_t track_base, hpfw_db **out) {
    if (!xs || !out || n < 0 || (n > 0 && !order)) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_xs_build_db: bad argument");
    std::vector<int64_t> src((size_t)n), len((size_t)n);
    for (int i = 0; i < n; ++i) {
        if (order[i] < 0 || order[i] >= int(xs->tracks.size())) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_xs_build_db: no track %d", order[i]);
        const XsTrack &t = xs->tracks[size_t(order[i])];
        if (t.hp_off < 0) HPFW_FAIL(HPFW_ERR_STATE, "hpfw_xs_build_db: track %d has not been hashed", order[i]);
        src[size_t(i)] = t.hp_off;
        len[size_t(i)] = t.words;
    }
    return hpfw_db_build_gather_device(xs->ctx, xs->hp.as<uint64_t>(), src.data(), len.data(), n, track_base, xs->ctx->stream, out);
}

int hpfw_xs_match(hpfw_xs *xs, hpfw_db *db, int topk, hpfw_match *out) {
    if (!xs || !db || !out) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_xs_match: NULL argument");
    const int n = int(xs->tracks.size());
    std::vector<int64_t> qo(size_t(n) + 1, 0);
    for (int t = 0; t < n; ++t) {
        if (xs->tracks[size_t(t)].hp_off != qo[size_t(t)])
            HPFW_FAIL(HPFW_ERR_STATE, "hpfw_xs_match: track %d has not been hashed (call hpfw_xs_hash_kept first)", t);
        qo[size_t(t) + 1] = qo[size_t(t)] + xs->tracks[size_t(t)].words;
    }
    return hpfw_db_find_topk_device(db, xs->hp.as<uint64_t>(), qo.data(), n, topk, out, xs->ctx->stream);
}

}  // extern "C"
