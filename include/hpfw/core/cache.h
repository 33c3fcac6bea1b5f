// include/hpfw/core/cache.h — cache::DriveCache, file-compatible with the reference's cereal archives
// (/root/reference/include/hpfw/core/cache.h:19-92): <dir>/accum_cov.cereal, <dir>/filters.cereal, <dir>/spectros/<stem>.
// An existing hpfw cache directory can therefore be served by the GPU path and vice versa.
#pragma once

#include <filesystem>
#include <string>
#include <utility>
#include <vector>

#include "../io/cereal_compat.h"
#include "../utils.h"

namespace hpfw::cache {

template <typename Algo>
class DriveCache {
public:
    explicit DriveCache(std::string cache) : cache_dir(std::move(cache)) {
        std::filesystem::create_directories(cache_dir + "spectros");
    }

    void set_spectro(const std::string &filename, const typename Algo::Spectrogram &f) const {
        io::save_matrix(cache_dir + "spectros/" + std::filesystem::path(filename).stem().string(), f);
    }
    /// The same file from a raw column-major float[121 x cols] buffer (the cache-writer threads of ParallelCollector hand over
    /// the pinned buffer the spectrogram was copied into; no Matrix copy in between).
    void set_spectro_raw(const std::string &filename, const float *data, int rows, int cols) const {
        io::save_matrix_raw(cache_dir + "spectros/" + std::filesystem::path(filename).stem().string(), data, rows, cols);
    }
    std::string spectro_path(const std::string &filename) const {
        return cache_dir + "spectros/" + std::filesystem::path(filename).stem().string();
    }
    void set_cov(const typename Algo::CovarianceMatrix &accum_cov) const {
        io::save_matrix(cache_dir + "accum_cov.cereal", accum_cov);
    }
    void set_filters(const typename Algo::Filters &f) const { io::save_matrix(cache_dir + "filters.cereal", f); }

    /// (path, spectrogram) of every file under <dir>/spectros/, loaded eagerly one at a time by the caller's loop.
    std::vector<std::string> spectro_files() const { return utils::get_dir_files(cache_dir + "spectros/"); }
    static std::pair<std::string, typename Algo::Spectrogram> load_spectro(const std::string &f) {
        typename Algo::Spectrogram s;
        io::load_matrix(f, s);
        return {f, std::move(s)};
    }
    bool get_cov(typename Algo::CovarianceMatrix &accum_cov) const {
        return io::load_matrix(cache_dir + "accum_cov.cereal", accum_cov);
    }
    bool get_filters(typename Algo::Filters &filters) const {
        return io::load_matrix(cache_dir + "filters.cereal", filters);
    }
    uint64_t size() const { return utils::count_dir_files(cache_dir + "spectros/"); }

private:
    const std::string cache_dir;
};

}  // namespace hpfw::cache
