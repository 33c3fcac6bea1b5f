// oracle/ref_build/ref_driver.cpp
//
// TEST INFRASTRUCTURE ONLY — never linked into or called by the product (hpfw_b200/).
//
// Thin extern "C" driver that instantiates the reference's OWN headers, compiled where they lie under
// /root/reference (nothing is copied), so tests and the fixture generator can call the real reference code:
//   include/hpfw/core/hashprint_handle.h      HashprintHandle<uint64_t, SH, 20, 80> static functions
//   include/hpfw/core/parallel_collector.h    ParallelCollector<Algo, Cache>::prepare / calc_hashprint
//   include/hpfw/spectrum/convert.h           amplitude_to_db / power_to_db
//   include/hpfw/audioproblems/live-song-id/storage.h   db::MemoryStorage<Collector>::build / find
// The spectrogram handler and the cache are template plug-in points of the reference; here they are an
// in-memory registry (essentia, cereal, boost are not installed, so cqt.h / cache.h cannot be compiled).
//
// Built by oracle/ref_build/Makefile into oracle/_ref/libhpfw_ref.so (git-ignored).

#include <atomic>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <mutex>
#include <optional>
#include <string>
#include <thread>
#include <vector>

#include <Eigen/Dense>
#include <Eigen/Eigenvalues>

#include "hpfw/core/hashprint_handle.h"
#include "hpfw/core/parallel_collector.h"
#include "hpfw/spectrum/convert.h"
#include "hpfw/audioproblems/live-song-id/storage.h"

namespace {

constexpr int kBins = 121;
using SpectroMat = Eigen::Matrix<float, kBins, Eigen::Dynamic>;

std::mutex g_mtx;
std::map<std::string, SpectroMat> g_spectros;  // "file name" -> spectrogram, filled by ref_register_spectrogram

// SpectrogramHandler plug-in (the reference's template parameter, hashprint_handle.h:51,56,72).
struct RegistrySH {
    using Spectrogram = SpectroMat;
    static Spectrogram spectrogram(const std::string &filename) {
        std::scoped_lock l(g_mtx);
        auto it = g_spectros.find(filename);
        if (it == g_spectros.end()) throw std::runtime_error("no such synthetic spectrogram: " + filename);
        return it->second;
    }
};

using Algo = hpfw::HashprintHandle<uint64_t, RegistrySH, 20, 80>;

// Cache plug-in (the reference's template-template parameter, parallel_collector.h:16,79); same member set as
// cache::DriveCache (cache.h:30-62) but held in memory.
template <typename A>
class MemCache {
public:
    explicit MemCache(std::string) {}
    void set_spectro(const std::string &filename, const typename A::Spectrogram &s) const {
        std::scoped_lock l(m);
        spectros.emplace_back(filename, s);
    }
    void set_cov(const typename A::CovarianceMatrix &c) const { cov = c; have_cov = true; }
    void set_filters(const typename A::Filters &f) const { filters = f; have_filters = true; }
    const std::vector<std::pair<std::string, typename A::Spectrogram>> &get_spectros() const { return spectros; }
    void get_cov(typename A::CovarianceMatrix &c) const { if (have_cov) c = cov; }
    void get_filters(typename A::Filters &f) const { if (have_filters) f = filters; }
    uint64_t size() const { return spectros.size(); }

    mutable std::mutex m;
    mutable std::vector<std::pair<std::string, typename A::Spectrogram>> spectros;
    mutable typename A::CovarianceMatrix cov;
    mutable typename A::Filters filters;
    mutable bool have_cov = false, have_filters = false;
};

using Collector = hpfw::ParallelCollector<Algo, MemCache>;
using Storage = hpfw::db::MemoryStorage<Collector>;

}  // namespace

extern "C" {

// ---- stage a3: convert.h:18-25 -----------------------------------------------------------------------------------
// in/out: column-major float[121 x cols] amplitudes -> dB
void ref_amplitude_to_db(float *spectro, int cols) {
    SpectroMat m = Eigen::Map<SpectroMat>(spectro, kBins, cols);
    SpectroMat out = hpfw::spectrum::amplitude_to_db(m);
    std::memcpy(spectro, out.data(), sizeof(float) * size_t(kBins) * size_t(cols));
}

// ---- stage a4: hashprint_handle.h:79-93 ----------------------------------------------------------------------------
// frames_out: row-major float[2420 x (cols-19)]
int ref_calc_frames(const float *spectro, int cols, float *frames_out) {
    SpectroMat m = Eigen::Map<const SpectroMat>(spectro, kBins, cols);
    Algo::Frames fr = Algo::calc_frames(m);
    std::memcpy(frames_out, fr.data(), sizeof(float) * size_t(fr.rows()) * size_t(fr.cols()));
    return int(fr.cols());
}

// ---- stages a4+a5: projection y = filters * frames (parallel_collector.h:57) ------------------------------------------
// filters: column-major float[64 x 2420]; y_out: column-major float[64 x (cols-19)]
int ref_project(const float *spectro, int cols, const float *filters, float *y_out) {
    SpectroMat m = Eigen::Map<const SpectroMat>(spectro, kBins, cols);
    Algo::Filters f = Eigen::Map<const Algo::Filters>(filters, 64, 2420);
    Algo::Frames fr = Algo::calc_frames(m);
    Eigen::Matrix<float, 64, Eigen::Dynamic> y = f * fr;
    std::memcpy(y_out, y.data(), sizeof(float) * 64 * size_t(y.cols()));
    return int(y.cols());
}

// ---- stages a4..a7 exactly as ParallelCollector::calc_hashprint does after the spectrogram (parallel_collector.h:56-58)
// returns number of hashprint words (= cols - 99) written to hp_out
int ref_hashprint_from_spectrogram(const float *spectro, int cols, const float *filters, uint64_t *hp_out) {
    SpectroMat m = Eigen::Map<const SpectroMat>(spectro, kBins, cols);
    Algo::Filters f = Eigen::Map<const Algo::Filters>(filters, 64, 2420);
    const Algo::Frames fr = Algo::calc_frames(m);
    const Algo::Fingerprint fp = Algo::calc_fingerprint(f * fr);
    const Algo::Hashprint hp = Algo::fingerprint_to_hashprint(fp);
    std::memcpy(hp_out, hp.data(), sizeof(uint64_t) * hp.size());
    return int(hp.size());
}

// ---- a6/a7 on an explicit y (column-major float[64 x ycols]) ------------------------------------------------------------
int ref_fingerprint_pack(const float *y, int ycols, uint64_t *hp_out) {
    Eigen::Matrix<float, 64, Eigen::Dynamic> ym = Eigen::Map<const Eigen::Matrix<float, 64, Eigen::Dynamic>>(y, 64, ycols);
    const Algo::Fingerprint fp = Algo::calc_fingerprint(ym);
    const Algo::Hashprint hp = Algo::fingerprint_to_hashprint(fp);
    std::memcpy(hp_out, hp.data(), sizeof(uint64_t) * hp.size());
    return int(hp.size());
}

// ---- a10: hashprint_handle.h:96-112 -------------------------------------------------------------------------------------
// cov_out: column-major float[2420 x 2420], covariance of this spectrogram's frames (calc_cov(frames^T))
void ref_calc_cov(const float *spectro, int cols, float *cov_out) {
    SpectroMat m = Eigen::Map<const SpectroMat>(spectro, kBins, cols);
    const Algo::Frames fr = Algo::calc_frames(m);
    const Algo::CovarianceMatrix c = Algo::calc_cov(fr.transpose());
    std::memcpy(cov_out, c.data(), sizeof(float) * size_t(c.rows()) * size_t(c.cols()));
}

// generic small-size variants for unit tests (rows x cols col-major in, cols x cols col-major out)
void ref_calc_cov_generic(const float *mat, int rows, int cols, float *cov_out) {
    Eigen::MatrixXf m = Eigen::Map<const Eigen::MatrixXf>(mat, rows, cols);
    const Algo::CovarianceMatrix c = Algo::calc_cov(m);
    std::memcpy(cov_out, c.data(), sizeof(float) * size_t(c.rows()) * size_t(c.cols()));
}

// cov: column-major float[2420 x 2420]; filters_out: column-major float[64 x 2420]
void ref_calc_filters(const float *cov, float *filters_out) {
    Algo::CovarianceMatrix c = Eigen::Map<const Eigen::MatrixXf>(cov, 2420, 2420);
    const Algo::Filters f = Algo::calc_filters(c);
    std::memcpy(filters_out, f.data(), sizeof(float) * 64 * 2420);
}

// ---- collector: ParallelCollector<Algo, MemCache> --------------------------------------------------------------------------
void ref_register_spectrogram(const char *name, const float *spectro, int cols) {
    std::scoped_lock l(g_mtx);
    g_spectros[name] = Eigen::Map<const SpectroMat>(spectro, kBins, cols);
}
void ref_clear_spectrograms() {
    std::scoped_lock l(g_mtx);
    g_spectros.clear();
}

void *ref_collector_new() { return new Collector(); }
void ref_collector_del(void *c) { delete static_cast<Collector *>(c); }

struct RefPrepared {
    std::vector<std::string> names;
    std::vector<std::vector<uint64_t>> hps;
};

// prepare(files): learns filters from the registered spectrograms, returns {stem, hashprint} list (unordered, as the reference)
void *ref_collector_prepare(void *c, const char **names, int n) {
    std::vector<std::string> files(names, names + n);
    auto res = static_cast<Collector *>(c)->prepare(files);
    auto *out = new RefPrepared();
    for (auto &p : res) {
        out->names.push_back(p.filename);
        out->hps.push_back(p.fingerprint);
    }
    return out;
}
int ref_prepared_count(void *p) { return int(static_cast<RefPrepared *>(p)->names.size()); }
const char *ref_prepared_name(void *p, int i) { return static_cast<RefPrepared *>(p)->names[i].c_str(); }
int ref_prepared_size(void *p, int i) { return int(static_cast<RefPrepared *>(p)->hps[i].size()); }
const uint64_t *ref_prepared_words(void *p, int i) { return static_cast<RefPrepared *>(p)->hps[i].data(); }
void ref_prepared_free(void *p) { delete static_cast<RefPrepared *>(p); }

// calc_hashprint(file) with the collector's current filters; returns word count, -1 on exception
int ref_collector_calc_hashprint(void *c, const char *name, uint64_t *out, int cap) {
    try {
        auto hp = static_cast<Collector *>(c)->calc_hashprint(name);
        if (int(hp.size()) > cap) return -2;
        std::memcpy(out, hp.data(), sizeof(uint64_t) * hp.size());
        return int(hp.size());
    } catch (const std::exception &) {
        return -1;
    }
}

// ---- matcher: db::MemoryStorage<Collector>::build / find (storage.h:21-64) ----------------------------------------------------
// The DB is given as a concatenated word array with offsets[R+1]; the track "filename" is its decimal index.
void *ref_storage_build(const uint64_t *words, const int64_t *offsets, int n_tracks) {
    tbb::concurrent_vector<Collector::FilenameFingerprintPair> v;
    for (int r = 0; r < n_tracks; ++r) {
        Collector::FilenameFingerprintPair p;
        p.filename = std::to_string(r);
        p.fingerprint.assign(words + offsets[r], words + offsets[r + 1]);
        v.push_back(std::move(p));
    }
    auto *s = new Storage();
    s->build(std::move(v));
    return s;
}
void ref_storage_del(void *s) { delete static_cast<Storage *>(s); }

// returns track index (-1 if the DB is empty: filename ""), cnt and offset exactly as SearchResult
int64_t ref_storage_find(void *s, const uint64_t *q, int k, uint64_t *cnt, int64_t *offset) {
    std::vector<uint64_t> hp(q, q + k);
    auto res = static_cast<const Storage *>(s)->find(hp);
    *cnt = uint64_t(res.cnt);
    *offset = res.offset;
    if (res.filename.empty()) return -1;
    return std::stoll(res.filename);
}

// many queries, thread-parallel over queries ("notebook style", liveid.ipynb:930-931); used as the CPU baseline.
// queries: concatenated words with qoffsets[Q+1]
void ref_storage_find_batch(void *s, const uint64_t *qwords, const int64_t *qoffsets, int n_queries,
                            int64_t *track_out, uint64_t *cnt_out, int64_t *offset_out, int n_threads) {
    const Storage *st = static_cast<const Storage *>(s);
    std::atomic<int> next{0};
    auto work = [&]() {
        for (int i = next.fetch_add(1); i < n_queries; i = next.fetch_add(1)) {
            std::vector<uint64_t> hp(qwords + qoffsets[i], qwords + qoffsets[i + 1]);
            auto res = st->find(hp);
            cnt_out[i] = uint64_t(res.cnt);
            offset_out[i] = res.offset;
            track_out[i] = res.filename.empty() ? -1 : std::stoll(res.filename);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
}

}  // extern "C"
