// include/hpfw/core/parallel_collector.h — ParallelCollector: per-file pipeline driver on the GPU.
//
// Same interface as /root/reference/include/hpfw/core/parallel_collector.h:16-139: prepare(files), calc_hashprint(file),
// save(), load(), FilenameFingerprintPair; `Algo` and `Cache` stay template plug-in points. What runs where:
//   calc_hashprint(file)  = decode (host) -> CQT (GPU) -> projection/threshold/pack (GPU); reference :54-59
//   prepare(files)        = per file: spectrogram (GPU), cache it (Cache::set_spectro, like :99), then hashprints of EVERY
//                           cached spectrogram (collect_fingerprints, :115-137). Per-file errors are caught and logged and
//                           the file is skipped (:101-103).
// Filter learning follows the reference too (:92-97,:111): every preprocessed spectrogram adds its frame covariance to
// accum_cov (kept in HBM while prepare() runs; structured GEMM in learn.cu instead of the 2420 x 2420 x frames SYRK) and
// filters = calc_filters(accum_cov) afterwards (GPU subspace iteration). As in the reference, accum_cov persists through
// save()/load(), so it is cumulative across runs. calc_hashprint() before any filters exist throws
// hpfw::Error(HPFW_ERR_STATE) (the reference would multiply by uninitialised memory).
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
        preprocess(filenames);
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

    /// Spectrograms, covariance accumulation, filters (reference :82-112).
    void preprocess(const std::vector<std::string> &filenames) {
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
                cache.set_spectro(filename, spectro);
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
        {
            std::scoped_lock l(ctx->mutex());
            device::check(hpfw_cov_get(ctx->get(), accum_cov.data()));
            // the reference divides by cache.size()+1 first (:111); a positive scale does not change the eigenvectors
            device::check(hpfw_calc_filters(ctx->get(), nullptr, f.data(), nullptr));
        }
        have_cov = true;
        set_filters(f);
    }

    void require_filters() const {
        if (!have_filters)
            throw Error(HPFW_ERR_STATE,
                        "no filters: prepare() has not run, load() found no cache/filters.cereal and set_filters() was "
                        "not called");
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
