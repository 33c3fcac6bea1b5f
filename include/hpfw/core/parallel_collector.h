// include/hpfw/core/parallel_collector.h — ParallelCollector: per-file pipeline driver on the GPU.
//
// Same interface as /root/reference/include/hpfw/core/parallel_collector.h:16-139: prepare(files), calc_hashprint(file),
// save(), load(), FilenameFingerprintPair; `Algo` and `Cache` stay template plug-in points. What runs where:
//   calc_hashprint(file)  = decode (host) -> CQT (GPU) -> projection/threshold/pack (GPU); reference :54-59
//   prepare(files)        = per file: spectrogram (GPU), cache it (Cache::set_spectro, like :99), then hashprints of EVERY
//                           cached spectrogram (collect_fingerprints, :115-137). Per-file errors are caught and logged and
//                           the file is skipped (:101-103).
// Filter learning (calc_cov / calc_filters, :92-97,:111) is index-time work outside this round's kernels:
// prepare() uses the filters that load() found in the cache (cache/filters.cereal written by hpfw itself or by
// set_filters()) and throws hpfw::Error(HPFW_ERR_STATE) when there are none.
#pragma once

#include <filesystem>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "../device.h"

namespace hpfw {

template <typename Algo, template <typename> typename Cache>
class ParallelCollector {
public:
    using Spectrogram = typename Algo::Spectrogram;
    using Frames = typename Algo::Frames;
    using CovarianceMatrix = typename Algo::CovarianceMatrix;
    using Filters = typename Algo::Filters;
    using Fingerprint = typename Algo::Fingerprint;
    using Hashprint = typename Algo::Hashprint;

    struct FilenameFingerprintPair {     // parallel_collector.h:26-35
        std::string filename;
        Hashprint fingerprint;
    };

    explicit ParallelCollector(const std::string &cache_dir = "cache/", int device = 0)
        : algo(), cache(cache_dir), ctx(device::Context::shared(device)) {
        filters.resize(Algo::NumOfFilters, Algo::FrameSize);
    }

    /// Process audio files and return {stem, hashprint} for every spectrogram in the cache (reference :48-52).
    std::vector<FilenameFingerprintPair> prepare(const std::vector<std::string> &filenames) {
        require_filters();
        for (const auto &filename : filenames) {
            try {
                cache.set_spectro(filename, algo.sh.spectrogram(filename));
            } catch (const std::exception &e) {
                std::cerr << "[hpfw] Error preprocessing '" << filename << "': " << e.what() << std::endl;
            }
        }
        save();
        return collect_fingerprints();
    }

    /// Query path (reference :54-59).
    Hashprint calc_hashprint(const std::string &filename) const {
        require_filters();
        return Algo::calc_hashprint(*ctx, algo.sh.spectrogram(filename));
    }

    /// Same from a decoded mono buffer: CQT, projection and packing without leaving the GPU.
    Hashprint calc_hashprint(const float *audio, int64_t n_samples) const {
        require_filters();
        const int n = hpfw_hashprint_words_for_samples(n_samples);
        Hashprint hp(n > 0 ? n : 0);
        int got = 0;
        std::scoped_lock l(ctx->mutex());
        device::check(hpfw_calc_hashprint_audio(ctx->get(), audio, n_samples, hp.data(), &got));
        return hp;
    }

    void save() const {      // reference :61-66
        if (have_cov) cache.set_cov(accum_cov);
        if (have_filters) cache.set_filters(filters);
    }

    void load() {            // reference :68-73
        have_cov = cache.get_cov(accum_cov);
        Filters f;
        if (cache.get_filters(f)) set_filters(f);
    }

    /// Install filters learned elsewhere (64 x 2420, column-major).
    void set_filters(const Filters &f) {
        Algo::set_filters(*ctx, f);
        filters = f;
        have_filters = true;
    }
    const Filters &get_filters() const { return filters; }
    device::Context &context() const { return *ctx; }

private:
    const Algo algo;
    CovarianceMatrix accum_cov;
    Filters filters;
    Cache<Algo> cache;
    std::shared_ptr<device::Context> ctx;
    bool have_filters = false, have_cov = false;

    void require_filters() const {
        if (!have_filters)
            throw Error(HPFW_ERR_STATE,
                        "no filters: load() found no cache/filters.cereal and set_filters() was not called "
                        "(GPU filter learning is not part of this build)");
        // another collector may have re-programmed the shared context
        Algo::set_filters(*ctx, filters);
    }

    std::vector<FilenameFingerprintPair> collect_fingerprints() const {   // reference :115-137
        std::vector<FilenameFingerprintPair> out;
        for (const auto &path : cache.spectro_files()) {
            try {
                auto p = Cache<Algo>::load_spectro(path);
                out.push_back({std::filesystem::path(p.first).stem().string(), Algo::calc_hashprint(*ctx, p.second)});
            } catch (const std::exception &e) {
                std::cerr << "[hpfw] Error fingerprinting '" << path << "': " << e.what() << std::endl;
            }
        }
        return out;
    }
};

}  // namespace hpfw
