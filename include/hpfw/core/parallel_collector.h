// include/hpfw/core/parallel_collector.h — ParallelCollector: per-file pipeline driver on the GPU.
//
// Same interface as /root/reference/include/hpfw/core/parallel_collector.h:16-139: prepare(files), calc_hashprint(file),
// save(), load(), FilenameFingerprintPair; `Algo` and `Cache` stay template plug-in points.
//
// The reference fans out over files with a task pool (:85-108 preprocess, :119-134 collect_fingerprints) and passes every
// intermediate through host memory and the cache directory. Here the same two phases are ONE device-resident pipeline
// (hpfw_xs_*, hpfw_b200/csrc/xstream.cu) whenever the SpectrogramHandler offers the `decode` hook (spectrum::CQT does):
//
//   prepare(files)   N decode threads read each file straight into a pinned staging slot -> H2D -> CQT on 4 lane streams
//                    -> the dB spectrogram STAYS in HBM; its frame covariance is added to accum_cov on a side stream (:92-97);
//                    cache/spectros/<stem> is written by background threads from a device->host copy, off the critical path
//                    (:99; they start when the pipeline has drained and finish after prepare() has returned); filters = calc_filters(accum_cov) (:111); then ONE batched projection/threshold/pack over all
//                    resident spectrograms plus every older file in cache/spectros/ (:115-137). prepare_device() leaves the
//                    hashprints in HBM for Storage::build_device; prepare() copies them out as the reference's pair list.
//   calc_hashprints_device(files)   the query side of the same pipeline (search(): all query files in one batch).
//   calc_hashprint(file)            the reference's single-file call (:54-59).
// A SpectrogramHandler without `decode` gets the reference's per-file flow (spectrogram(filename) on the host side, one
// file at a time). Per-file errors are caught, logged and the file is skipped (:101-103). As in the reference, accum_cov
// persists through save()/load(), so it is cumulative across runs. calc_hashprint() before any filters exist throws
// hpfw::Error(HPFW_ERR_STATE) (the reference would multiply by uninitialised memory).
#pragma once

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <filesystem>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <type_traits>
#include <utility>
#include <vector>

#include "../device.h"
#include "../io/cereal_compat.h"

namespace hpfw {

namespace detail {
template <typename SH, typename = void>
struct has_decode : std::false_type {};
template <typename SH>
struct has_decode<SH, std::void_t<decltype(SH::decode(std::declval<const std::string &>(),
                                                       std::declval<void *(*)(size_t)>()))>> : std::true_type {};
template <typename C, typename = void>
struct has_raw_cache : std::false_type {};
template <typename C>
struct has_raw_cache<C, std::void_t<decltype(std::declval<const C &>().set_spectro_raw(
                            std::declval<const std::string &>(), std::declval<const float *>(), 0, 0))>> : std::true_type {};

/// HPFW_TRACE=1: phase timings of prepare() / search batches on stderr
struct PhaseTrace {
    bool on = std::getenv("HPFW_TRACE") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void mark(const char *what) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        std::cerr << "[hpfw trace] " << what << ": " << std::chrono::duration<double, std::milli>(t1 - t0).count() << " ms"
                  << std::endl;
        t0 = t1;
    }
};

inline unsigned worker_count(size_t jobs) {
    unsigned hw = std::thread::hardware_concurrency();
    if (const char *env = std::getenv("HPFW_DECODE_THREADS")) hw = static_cast<unsigned>(std::max(1, std::atoi(env)));
    if (hw == 0) hw = 4;
    return static_cast<unsigned>(std::max<size_t>(1, std::min<size_t>(hw, jobs)));
}
}  // namespace detail

template <typename Algo, template <typename> typename Cache>
class ParallelCollector {
public:
    using Spectrogram = typename Algo::Spectrogram;
    using Frames = typename Algo::Frames;
    using CovarianceMatrix = typename Algo::CovarianceMatrix;
    using Filters = typename Algo::Filters;
    using Fingerprint = typename Algo::Fingerprint;
    using Hashprint = typename Algo::Hashprint;
    using SpectrogramHandler = std::decay_t<decltype(std::declval<Algo>().sh)>;
    static constexpr bool batched = detail::has_decode<SpectrogramHandler>::value;

    struct FilenameFingerprintPair {     // parallel_collector.h:26-35
        std::string filename;
        Hashprint fingerprint;
    };

    /// Hashprints that are still in HBM: `names[i]` belongs to track `order[i]` of the extraction stream `xs`.
    /// prepare_device() returns the database side (DB order = order of names), calc_hashprints_device() the query side.
    struct DeviceHashprints {
        std::shared_ptr<device::ExtractionStream> xs;
        std::shared_ptr<device::Context> ctx;
        std::vector<std::string> names;
        std::vector<int> order;
        std::vector<int> words;
        std::vector<std::string> errors;     // query side: message per input file ("" = ok); names/order hold the ok ones
        std::vector<int> slot_of_input;      // query side: index into names/order per input file, -1 = failed
        uint64_t generation = 0;             // of xs when this was produced: the next prepare / search batch invalidates it
        bool valid() const { return xs && xs->generation == generation; }
    };

    explicit ParallelCollector(const std::string &cache_dir = "cache/", int device = 0)
        : algo(), cache(std::make_unique<Cache<Algo>>(cache_dir)), ctx(device::Context::shared(device)), device_(device) {
        filters.resize(Algo::NumOfFilters, Algo::FrameSize);
        static std::atomic<uint64_t> next_id{1};
        id_ = next_id.fetch_add(1) << 32;
        if (const char *env = std::getenv("HPFW_CACHE_SPECTROGRAMS")) cache_spectrograms = std::atoi(env) != 0;
    }
    ~ParallelCollector() {
        try { flush_cache_writes(); } catch (...) {}
    }

    /// Process audio files and return {stem, hashprint} for every spectrogram in the cache (reference :48-52).
    std::vector<FilenameFingerprintPair> prepare(const std::vector<std::string> &filenames) {
        if constexpr (batched) {
            DeviceHashprints d = prepare_device(filenames);
            std::vector<FilenameFingerprintPair> out(d.names.size());
            std::scoped_lock l(ctx->mutex());
            // one device->host copy of the whole hashprint store, then one slice per track
            const int nt = hpfw_xs_tracks(d.xs->get());
            std::vector<int64_t> off(static_cast<size_t>(nt)), len(static_cast<size_t>(nt));
            device::check(hpfw_xs_hashprints_device(d.xs->get(), nullptr, off.data(), len.data()));
            int64_t total = 0;
            for (int t = 0; t < nt; ++t)
                if (off[static_cast<size_t>(t)] >= 0) total = std::max(total, off[static_cast<size_t>(t)] + len[static_cast<size_t>(t)]);
            std::vector<uint64_t> store(static_cast<size_t>(total));
            device::check(hpfw_xs_hashprints_host(d.xs->get(), store.data(), total));
            for (size_t i = 0; i < d.names.size(); ++i) {
                const size_t t = static_cast<size_t>(d.order[i]);
                out[i].filename = d.names[i];
                out[i].fingerprint.assign(store.begin() + off[t], store.begin() + off[t] + len[t]);
            }
            return out;
        } else {
            preprocess_serial(filenames);
            save();
            return collect_fingerprints_serial();
        }
    }

    /// prepare() with the result left on the GPU (see DeviceHashprints). The spectrograms stay resident until the background
    /// cache writers have finished (flush_cache_writes(), called by the next prepare / the destructor).
    /// start_writers = false leaves the cache/spectros writers parked until begin_cache_writes() (or any flush): index() builds
    /// its storage first, so that the writers' device->host copies do not sit between the storage's allocations and kernels.
    template <bool B = batched, typename = std::enable_if_t<B>>
    DeviceHashprints prepare_device(const std::vector<std::string> &filenames, bool start_writers = true) {
        detail::PhaseTrace trace;
        flush_cache_writes();
        std::unique_lock<std::mutex> l(ctx->mutex());
        if (have_cov) device::check(hpfw_cov_set(ctx->get(), accum_cov.data()));
        else device::check(hpfw_cov_reset(ctx->get()));
        if (!ixs) ixs = std::make_shared<device::ExtractionStream>(*ctx, slots_for(filenames.size()), size_t(16) << 20);
        device::check(hpfw_xs_reset(ixs->get()));
        ++ixs->generation;
        hpfw_xs *xs = ixs->get();

        // ---- phase 1 (reference :82-108): decode || upload || CQT || covariance, spectrograms stay in HBM
        std::map<std::string, int> track_of_stem;          // this run's stems -> stream track (resident or spilled: -1)
        size_t added = 0, resident = 0;
        run_decoders(
            xs, filenames,
            [&](size_t i, int slot, const auto &info) -> bool {
                int track = -1;
                const int flags = (info.pcm16 ? HPFW_XS_PCM16 : 0) | HPFW_XS_COV;
                int st = hpfw_xs_submit(xs, slot, info.n_samples, flags, &track);
                if (st == HPFW_ERR_LIMIT) {
                    // HBM budget for resident spectrograms reached: spill what is resident through the cache files (as the
                    // reference always does) and decode this file again
                    if (!cache_spectrograms)
                        throw Error(st, std::string("hpfw_b200: ") + hpfw_last_error() + " (spectrogram cache is disabled)");
                    if (resident == 0)
                        throw Error(st, std::string("hpfw_b200: ") + hpfw_last_error() + " (one track alone exceeds the budget)");
                    resident = 0;
                    l.unlock();
                    flush_cache_writes();
                    l.lock();
                    device::check(hpfw_xs_drop_kept(xs));
                    for (auto &kv : track_of_stem) kv.second = -1;
                    return false;                           // retry
                }
                if (st != HPFW_OK) {                          // this file only (too short, bad length): log and skip (:101-103)
                    std::cerr << "[hpfw] Error preprocessing '" << filenames[i] << "': " << hpfw_last_error() << std::endl;
                    return true;
                }
                const std::string stem = std::filesystem::path(filenames[i]).stem().string();
                track_of_stem[stem] = track;
                ++added;
                ++resident;
                if (cache_spectrograms) enqueue_cache_write(filenames[i], track);
                return true;
            },
            [&](size_t i, const std::string &what) {
                std::cerr << "[hpfw] Error preprocessing '" << filenames[i] << "': " << what << std::endl;
            },
            l);
        trace.mark("prepare: decode + upload + CQT + covariance enqueued");
        device::check(hpfw_xs_wait(xs));
        trace.mark("prepare: pipeline drained");
        if (added == 0 && !have_cov) {
            if (!have_filters)
                throw Error(HPFW_ERR_STATE, "prepare(): no readable audio file and no cached covariance to learn filters from");
        } else {
            // filters = calc_filters(accum_cov / (cache.size()+1)) (:111); a positive scale does not change the eigenvectors
            accum_cov.resize(Algo::FrameSize, Algo::FrameSize);
            Filters f(Algo::NumOfFilters, Algo::FrameSize);
            device::check(hpfw_cov_get(ctx->get(), accum_cov.data()));
            device::check(hpfw_calc_filters(ctx->get(), nullptr, f.data(), nullptr));
            have_cov = true;
            install_filters_locked(f);
        }
        trace.mark("prepare: covariance fetched, filters learned");
        l.unlock();
        save();                                             // reference :50
        l.lock();
        trace.mark("prepare: save()");

        // ---- phase 2 (reference :115-137): hashprints of EVERY spectrogram in the cache: the resident ones from HBM, the rest
        // (earlier runs, or spilled above) from their files
        require_filters_locked();
        device::check(hpfw_xs_hash_kept(xs));
        std::vector<std::string> on_disk;
        for (const auto &path : cache->spectro_files()) {
            const std::string stem = std::filesystem::path(path).filename().string();
            if (stem.size() > 5 && stem.compare(stem.size() - 5, 5, ".tmp~") == 0) continue;
            auto it = track_of_stem.find(stem);
            if (it != track_of_stem.end() && it->second >= 0) continue;     // resident: already hashed from HBM
            on_disk.push_back(path);
        }
        if (!on_disk.empty()) {
            l.unlock();
            flush_cache_writes();                           // spilled files of this run must be complete on disk
            l.lock();
            load_cached_spectrograms(xs, on_disk, track_of_stem, l);
        }
        device::check(hpfw_xs_hash_kept(xs));

        trace.mark("prepare: hashed");
        if (start_writers) start_cache_writers();
        DeviceHashprints out;
        out.xs = ixs;
        out.ctx = ctx;
        out.generation = ixs->generation;
        for (const auto &kv : track_of_stem) {              // std::map: sorted by stem = sorted cache paths, as before
            if (kv.second < 0) continue;
            int cols = 0, words = 0;
            device::check(hpfw_xs_track_info(xs, kv.second, &cols, &words, nullptr));
            // the reference names a DB entry path(cache file).stem() (:126-128): a stem with a dot in it loses its tail
            out.names.push_back(std::filesystem::path(kv.first).stem().string());
            out.order.push_back(kv.second);
            out.words.push_back(words);
        }
        return out;
    }

    /// The query side of the pipeline: hashprints of all files in one batch, left in HBM (LiveSongIdentification::search).
    /// Per-file failures (unreadable, too short, longer than the matcher's query limit) are reported in `errors`.
    template <bool B = batched, typename = std::enable_if_t<B>>
    DeviceHashprints calc_hashprints_device(const std::vector<std::string> &filenames) {
        std::unique_lock<std::mutex> l(ctx->mutex());
        require_filters_locked();
        if (!qxs) qxs = std::make_shared<device::ExtractionStream>(*ctx, slots_for(filenames.size()), size_t(1) << 20);
        hpfw_xs *xs = qxs->get();
        device::check(hpfw_xs_reset(xs));
        ++qxs->generation;
        DeviceHashprints out;
        out.xs = qxs;
        out.ctx = ctx;
        out.generation = qxs->generation;
        out.errors.assign(filenames.size(), "");
        out.slot_of_input.assign(filenames.size(), -1);
        std::vector<std::pair<int, size_t>> by_track;        // (stream track, input index)
        run_decoders(
            xs, filenames,
            [&](size_t i, int slot, const auto &info) -> bool {
                const int words = hpfw_hashprint_words_for_samples(info.n_samples);
                if (words > HPFW_MAX_QUERY_WORDS) {
                    hpfw_xs_release(xs, slot);
                    out.errors[i] = "query of " + std::to_string(words) + " hashprint words exceeds the matcher's limit of " +
                                    std::to_string(HPFW_MAX_QUERY_WORDS) + " (about 64 s of audio)";
                    return true;
                }
                int track = -1;
                const int st = hpfw_xs_submit(xs, slot, info.n_samples, info.pcm16 ? HPFW_XS_PCM16 : 0, &track);
                if (st != HPFW_OK) out.errors[i] = std::string("hpfw_b200: ") + hpfw_last_error();
                else by_track.emplace_back(track, i);
                return true;
            },
            [&](size_t i, const std::string &what) { out.errors[i] = what; }, l);
        device::check(hpfw_xs_hash_kept(xs));
        std::sort(by_track.begin(), by_track.end());
        for (const auto &p : by_track) {
            int cols = 0, words = 0;
            device::check(hpfw_xs_track_info(xs, p.first, &cols, &words, nullptr));
            out.slot_of_input[p.second] = static_cast<int>(out.order.size());
            out.names.push_back(filenames[p.second]);
            out.order.push_back(p.first);
            out.words.push_back(words);
        }
        return out;
    }

    /// Query path (reference :54-59).
    Hashprint calc_hashprint(const std::string &filename) const {
        if constexpr (batched) {
            std::vector<unsigned char> buf;
            const auto info = SpectrogramHandler::decode(filename, [&](size_t bytes) -> void * {
                buf.resize(bytes);
                return buf.data();
            });
            std::scoped_lock l(ctx->mutex());
            require_filters_locked();
            const int n = hpfw_hashprint_words_for_samples(info.n_samples);
            Hashprint hp(n > 0 ? n : 0);
            int got = 0;
            if (info.pcm16)
                device::check(hpfw_calc_hashprint_pcm16(ctx->get(), reinterpret_cast<const int16_t *>(buf.data()), info.n_samples,
                                                        hp.data(), &got));
            else
                device::check(hpfw_calc_hashprint_audio(ctx->get(), reinterpret_cast<const float *>(buf.data()), info.n_samples,
                                                        hp.data(), &got));
            return hp;
        } else {
            const Spectrogram spectro = algo.sh.spectrogram(filename);
            std::scoped_lock l(ctx->mutex());
            require_filters_locked();
            return hashprint_locked(spectro);
        }
    }

    /// Same from a decoded mono buffer: CQT, projection and packing without leaving the GPU.
    Hashprint calc_hashprint(const float *audio, int64_t n_samples) const {
        const int n = hpfw_hashprint_words_for_samples(n_samples);
        Hashprint hp(n > 0 ? n : 0);
        int got = 0;
        std::scoped_lock l(ctx->mutex());
        require_filters_locked();
        device::check(hpfw_calc_hashprint_audio(ctx->get(), audio, n_samples, hp.data(), &got));
        return hp;
    }

    void save() const {      // reference :61-66
        if (have_cov) cache->set_cov(accum_cov);
        if (have_filters) cache->set_filters(filters);
    }

    void load() {            // reference :68-73
        have_cov = cache->get_cov(accum_cov);
        Filters f;
        if (cache->get_filters(f)) set_filters(f);
    }

    /// The cache directory save()/load() and the spectrogram cache use from now on (the learned state is kept: the reference's
    /// C wrapper ignores its `cache` argument, modules/python/parallel_collector_wrapper.cpp:56-62; here it is honoured).
    void set_cache_dir(const std::string &dir) {
        flush_cache_writes();
        cache = std::make_unique<Cache<Algo>>(dir);
    }

    /// Install filters learned elsewhere (64 x 2420, column-major).
    void set_filters(const Filters &f) {
        std::scoped_lock l(ctx->mutex());
        install_filters_locked(f);
    }
    const Filters &get_filters() const { return filters; }
    device::Context &context() const { return *ctx; }
    int device() const { return device_; }

    /// Whether prepare() writes cache/spectros/<stem> (the reference always does, :99). Off: the spectrograms only ever exist
    /// in HBM; a later run cannot re-hash them from the cache.
    void set_cache_spectrograms(bool on) { cache_spectrograms = on; }

    /// Start the background cache writers of the last prepare_device(files, false); a no-op when nothing is pending.
    void begin_cache_writes() { start_cache_writers(); }

    /// Wait for the background cache writers and release the resident spectrograms of the last prepare().
    void flush_cache_writes() {
        start_cache_writers();
        {
            std::unique_lock<std::mutex> wl(wq_m);
            wq_stop = true;
        }
        wq_cv.notify_all();
        for (auto &t : writers) t.join();
        writers.clear();
        wq_stop = false;
        if (ixs && writers_used) {
            std::scoped_lock l(ctx->mutex());
            hpfw_xs_drop_kept(ixs->get());
            writers_used = false;
        }
    }

private:
    Algo algo;
    CovarianceMatrix accum_cov;
    Filters filters;
    std::unique_ptr<Cache<Algo>> cache;       // re-created by set_cache_dir (a Cache plug-in only needs its (dir) constructor)
    std::shared_ptr<device::Context> ctx;
    int device_ = 0;
    bool have_filters = false, have_cov = false, cache_spectrograms = true;
    uint64_t id_ = 0, filters_gen = 0;
    std::shared_ptr<device::ExtractionStream> ixs, qxs;     // index side / query side

    // background cache writers
    struct WriteJob { std::string filename; int track; };
    std::vector<std::thread> writers;
    std::deque<WriteJob> wq;
    std::mutex wq_m;
    std::condition_variable wq_cv;
    bool wq_stop = false, writers_used = false;

    static int slots_for(size_t files) { return static_cast<int>(2 * detail::worker_count(files) + 2); }

    void install_filters_locked(const Filters &f) {
        if (f.rows() != static_cast<std::ptrdiff_t>(Algo::NumOfFilters) || f.cols() != static_cast<std::ptrdiff_t>(Algo::FrameSize))
            throw Error(HPFW_ERR_ARG, "filters must be 64 x 2420");
        device::check(hpfw_set_filters(ctx->get(), f.data()));
        filters = f;
        have_filters = true;
        ctx->filters_tag = id_ | (++filters_gen & 0xFFFFFFFFu);
    }

    /// The shared device context holds one filter set: re-upload only when another collector (or another generation of this
    /// one) installed the current one. Called with the context mutex held, which the caller keeps for the hashprint call.
    void require_filters_locked() const {
        if (!have_filters)
            throw Error(HPFW_ERR_STATE,
                        "no filters: prepare() has not run, load() found no cache/filters.cereal and set_filters() was "
                        "not called");
        const uint64_t tag = id_ | (filters_gen & 0xFFFFFFFFu);
        if (ctx->filters_tag != tag) {
            device::check(hpfw_set_filters(ctx->get(), filters.data()));
            ctx->filters_tag = tag;
        }
    }

    Hashprint hashprint_locked(const Spectrogram &spectro) const {
        const int cols = static_cast<int>(spectro.cols());
        const int n = hpfw_hashprint_words_for_cols(cols);
        Hashprint hp(n > 0 ? n : 0);
        int got = 0;
        device::check(hpfw_hashprint_from_spectrogram(ctx->get(), spectro.data(), cols, hp.data(), &got));
        return hp;
    }

    /// N decode threads (reference: tf::Taskflow::parallel_for over files, :85-108) fill pinned staging slots; THIS thread, which
    /// holds the context mutex `l`, submits them in arrival order. on_ready returns false to have the file decoded again.
    template <typename OnReady, typename OnError>
    void run_decoders(hpfw_xs *xs, const std::vector<std::string> &filenames, OnReady &&on_ready, OnError &&on_error,
                      std::unique_lock<std::mutex> & /*context lock held by the caller*/) {
        using DecodeInfo = decltype(SpectrogramHandler::decode(std::declval<const std::string &>(),
                                                               std::declval<void *(*)(size_t)>()));
        struct Ready { size_t i; int slot; DecodeInfo info; std::string error; };
        std::deque<Ready> ready;
        std::deque<size_t> todo;
        for (size_t i = 0; i < filenames.size(); ++i) todo.push_back(i);
        std::mutex qm;
        std::condition_variable qcv;
        size_t outstanding = filenames.size();   // files not yet consumed by the submitting thread
        bool stop = false;
        // HPFW_TRACE: where the wall time of this phase goes (summed over the decode threads / on the submitting thread)
        const bool tracing = std::getenv("HPFW_TRACE") != nullptr;
        std::atomic<int64_t> us_acquire{0}, us_decode{0};
        int64_t us_starved = 0, us_submit = 0;
        auto now_us = [] {
            return std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
        };
        auto worker = [&]() {
            for (;;) {
                size_t i;
                {
                    std::unique_lock<std::mutex> ql(qm);
                    qcv.wait(ql, [&] { return stop || !todo.empty(); });
                    if (todo.empty()) return;
                    i = todo.front();
                    todo.pop_front();
                }
                Ready r{i, -1, DecodeInfo(), ""};
                const int64_t t_dec = tracing ? now_us() : 0;
                int64_t t_acq = 0;
                try {
                    r.info = SpectrogramHandler::decode(filenames[i], [&](size_t bytes) -> void * {
                        void *p = nullptr;
                        const int64_t a0 = tracing ? now_us() : 0;
                        device::check(hpfw_xs_acquire(xs, bytes, &r.slot, &p));
                        if (tracing) t_acq += now_us() - a0;
                        return p;
                    });
                    if (tracing) {
                        us_acquire += t_acq;
                        us_decode += now_us() - t_dec - t_acq;
                    }
                } catch (const std::exception &e) {
                    r.error = e.what();
                    if (r.error.empty()) r.error = "decode failed";
                    if (r.slot >= 0) hpfw_xs_release(xs, r.slot);
                    r.slot = -1;
                }
                {
                    std::unique_lock<std::mutex> ql(qm);
                    ready.push_back(std::move(r));
                }
                qcv.notify_all();
            }
        };
        std::vector<std::thread> pool;
        const unsigned nthreads = detail::worker_count(filenames.size());
        for (unsigned t = 0; t < nthreads && !filenames.empty(); ++t) pool.emplace_back(worker);
        std::exception_ptr fatal;
        while (outstanding > 0) {
            Ready r;
            const int64_t t_wait = tracing ? now_us() : 0;
            {
                std::unique_lock<std::mutex> ql(qm);
                qcv.wait(ql, [&] { return !ready.empty(); });
                r = std::move(ready.front());
                ready.pop_front();
            }
            const int64_t t_got = tracing ? now_us() : 0;
            us_starved += t_got - t_wait;
            struct SubmitTimer {
                int64_t &acc, t0;
                bool on;
                std::function<int64_t()> clock;
                ~SubmitTimer() { if (on) acc += clock() - t0; }
            } submit_timer{us_submit, t_got, tracing, now_us};
            if (fatal) {                                   // draining after a fatal error: give the slots back
                if (r.slot >= 0) hpfw_xs_release(xs, r.slot);
                --outstanding;
                continue;
            }
            if (!r.error.empty()) {
                on_error(r.i, r.error);
                --outstanding;
                continue;
            }
            try {
                if (on_ready(r.i, r.slot, r.info)) {
                    --outstanding;
                } else {
                    {
                        std::unique_lock<std::mutex> ql(qm);
                        todo.push_back(r.i);
                    }
                    qcv.notify_all();
                }
            } catch (...) {
                fatal = std::current_exception();
                --outstanding;
                std::unique_lock<std::mutex> ql(qm);
                outstanding -= todo.size();                // files no decoder has started on are dropped
                todo.clear();
            }
        }
        {
            std::unique_lock<std::mutex> ql(qm);
            stop = true;
        }
        qcv.notify_all();
        for (auto &t : pool) t.join();
        if (tracing)
            std::cerr << "[hpfw trace] decoders: " << nthreads << " threads, per thread " << us_decode.load() / 1000 / std::max(1u, nthreads)
                      << " ms reading files + " << us_acquire.load() / 1000 / std::max(1u, nthreads)
                      << " ms waiting for a staging slot; submitting thread " << us_submit / 1000 << " ms in submit, "
                      << us_starved / 1000 << " ms waiting for a decoded file" << std::endl;
        if (fatal) std::rethrow_exception(fatal);
    }

    // ---- background cache writers: cache/spectros/<stem> from a device->host copy of the resident spectrogram (reference :99)
    void enqueue_cache_write(const std::string &filename, int track) {
        {
            std::unique_lock<std::mutex> wl(wq_m);
            wq.push_back({filename, track});
        }
        writers_used = true;
    }

    /// The writer threads start once the extraction pipeline has drained (end of prepare_device, or a flush): while the decode
    /// threads stream files at the host's memory bandwidth, the writers' device->host copies and file writes would compete with
    /// them (measured: index() 10 % slower). The spectrograms stay resident in HBM until the writers are done either way.
    void start_cache_writers() {
        size_t pending;
        {
            std::unique_lock<std::mutex> wl(wq_m);
            pending = wq.size();
        }
        while (pending > 0 && writers.size() < std::min<size_t>(4, pending)) writers.emplace_back([this] { writer_loop(); });
        wq_cv.notify_all();
    }

    void writer_loop() {
        device::PinnedBuffer pin;
        for (;;) {
            WriteJob job;
            {
                std::unique_lock<std::mutex> wl(wq_m);
                wq_cv.wait(wl, [&] { return wq_stop || !wq.empty(); });
                if (wq.empty()) return;
                job = std::move(wq.front());
                wq.pop_front();
            }
            try {
                int cols = 0, words = 0, resident = 0;
                device::check(hpfw_xs_track_info(ixs->get(), job.track, &cols, &words, &resident));
                float *buf = static_cast<float *>(pin.reserve(sizeof(float) * size_t(cols) * HPFW_BINS));
                device::check(hpfw_xs_fetch_spectrogram(ixs->get(), job.track, buf));
                if constexpr (detail::has_raw_cache<Cache<Algo>>::value) {
                    cache->set_spectro_raw(job.filename, buf, HPFW_BINS, cols);
                } else {
                    Spectrogram s(HPFW_BINS, cols);
                    std::copy(buf, buf + size_t(cols) * HPFW_BINS, s.data());
                    cache->set_spectro(job.filename, s);
                }
            } catch (const std::exception &e) {
                std::cerr << "[hpfw] Error caching the spectrogram of '" << job.filename << "': " << e.what() << std::endl;
            }
        }
    }

    /// Spectrogram files that are not resident (earlier runs; spilled): file -> pinned slot -> HBM, hashed in arena-sized batches.
    void load_cached_spectrograms(hpfw_xs *xs, const std::vector<std::string> &paths, std::map<std::string, int> &track_of_stem,
                                  std::unique_lock<std::mutex> & /*context lock held by the caller*/) {
        for (const auto &path : paths) {
            try {
                int32_t rows = 0, cols = 0;
                if (!io::matrix_header(path, rows, cols) || rows != HPFW_BINS)
                    throw std::runtime_error("'" + path + "': not a 121-row spectrogram file");
                for (int attempt = 0;; ++attempt) {
                    int slot = -1, track = -1;
                    void *p = nullptr;
                    device::check(hpfw_xs_acquire(xs, sizeof(float) * size_t(cols) * HPFW_BINS, &slot, &p));
                    try {
                        io::load_matrix_payload(path, static_cast<float *>(p), size_t(cols) * HPFW_BINS);
                    } catch (...) {
                        hpfw_xs_release(xs, slot);
                        throw;
                    }
                    const int st = hpfw_xs_submit_spectrogram(xs, slot, cols, 0, &track);
                    if (st == HPFW_ERR_LIMIT && attempt == 0) {    // arena full: hash what is resident, free it, go on
                        device::check(hpfw_xs_hash_kept(xs));
                        device::check(hpfw_xs_drop_kept(xs));
                        continue;
                    }
                    device::check(st);
                    track_of_stem[std::filesystem::path(path).filename().string()] = track;
                    break;
                }
            } catch (const std::exception &e) {
                std::cerr << "[hpfw] Error fingerprinting '" << path << "': " << e.what() << std::endl;
            }
        }
    }

    // ---- the reference's per-file flow for SpectrogramHandlers without the decode hook
    void preprocess_serial(const std::vector<std::string> &filenames) {
        size_t added = 0;
        {
            std::scoped_lock l(ctx->mutex());
            if (have_cov) device::check(hpfw_cov_set(ctx->get(), accum_cov.data()));
            else device::check(hpfw_cov_reset(ctx->get()));
        }
        for (const auto &filename : filenames) {
            try {
                const Spectrogram spectro = algo.sh.spectrogram(filename);
                {
                    std::scoped_lock l(ctx->mutex());
                    device::check(hpfw_cov_add_spectrogram(ctx->get(), spectro.data(), static_cast<int>(spectro.cols())));
                }
                cache->set_spectro(filename, spectro);
                ++added;
            } catch (const std::exception &e) {
                std::cerr << "[hpfw] Error preprocessing '" << filename << "': " << e.what() << std::endl;
            }
        }
        if (added == 0 && !have_cov) {
            if (have_filters) return;      // nothing new to learn from: keep the loaded filters
            throw Error(HPFW_ERR_STATE, "prepare(): no readable audio file and no cached covariance to learn filters from");
        }
        accum_cov.resize(Algo::FrameSize, Algo::FrameSize);
        Filters f(Algo::NumOfFilters, Algo::FrameSize);
        std::scoped_lock l(ctx->mutex());
        device::check(hpfw_cov_get(ctx->get(), accum_cov.data()));
        device::check(hpfw_calc_filters(ctx->get(), nullptr, f.data(), nullptr));
        have_cov = true;
        install_filters_locked(f);
    }

    std::vector<FilenameFingerprintPair> collect_fingerprints_serial() const {   // reference :115-137
        std::vector<FilenameFingerprintPair> out;
        for (const auto &path : cache->spectro_files()) {
            try {
                auto p = Cache<Algo>::load_spectro(path);
                std::scoped_lock l(ctx->mutex());
                require_filters_locked();
                out.push_back({std::filesystem::path(p.first).stem().string(), hashprint_locked(p.second)});
            } catch (const std::exception &e) {
                std::cerr << "[hpfw] Error fingerprinting '" << path << "': " << e.what() << std::endl;
            }
        }
        return out;
    }
};

}  // namespace hpfw
