// hpfw_b200/csrc/pyhpfw_abi.cu — the reference's par_collector_* C ABI (include/hpfw_b200_pyhpfw.h) over the C++ host
// classes in include/hpfw/. Host code only (compiled by nvcc with the rest of the library).
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "common.cuh"
#include "../../include/hpfw_b200_pyhpfw.h"
#include "../../include/hpfw/audioproblems/live-song-id/live_song_id.h"

using Collector = hpfw::DefaultLiveIdCollector;

struct LiveIdCollector {
    std::string cache_dir = "cache/";
    std::unique_ptr<Collector> impl;
    Collector &get() {
        if (!impl) impl.reset(new Collector(cache_dir));
        return *impl;
    }
    // The reference's wrapper ignores the `cache` argument of save/load and always uses the live collector
    // (modules/python/parallel_collector_wrapper.cpp:56-62). Here the directory is honoured, and the SAME collector keeps its
    // learned state (filters, accum_cov): prepare() -> save(dir) -> load(dir) -> calc_hashprint works.
    void retarget(const char *cache) {
        if (!cache || !cache[0]) return;
        std::string d(cache);
        if (d.back() != '/') d.push_back('/');
        if (d == cache_dir && impl) return;
        cache_dir = d;
        if (impl) impl->set_cache_dir(cache_dir);
        else impl.reset(new Collector(cache_dir));
    }
};

extern "C" {

LiveIdCollector *par_collector_new(void) { return new LiveIdCollector(); }

void par_collector_del(LiveIdCollector *c) { delete c; }

FilenameHashprintPair *par_collector_prepare(LiveIdCollector *c, const char **filenames, int n, int *got) {
    if (got) *got = 0;
    try {
        if (!c || (n > 0 && !filenames) || !got) throw std::invalid_argument("par_collector_prepare: NULL argument");
        auto res = c->get().prepare(std::vector<std::string>(filenames, filenames + n));
        auto *out = new FilenameHashprintPair[res.size() ? res.size() : 1];
        for (size_t i = 0; i < res.size(); ++i) {
            out[i].hp_size = static_cast<int>(res[i].fingerprint.size());
            out[i].filename = new char[res[i].filename.size() + 1];
            std::memcpy(out[i].filename, res[i].filename.c_str(), res[i].filename.size() + 1);
            out[i].hashprint = new uint64_t[res[i].fingerprint.size() ? res[i].fingerprint.size() : 1];
            std::memcpy(out[i].hashprint, res[i].fingerprint.data(), sizeof(uint64_t) * res[i].fingerprint.size());
        }
        *got = static_cast<int>(res.size());
        return out;
    } catch (const std::exception &e) {
        hpfw_b200::set_error("par_collector_prepare: %s", e.what());
        return nullptr;
    }
}

void prepare_result_free(FilenameHashprintPair *res, int got) {
    if (!res) return;
    for (int i = 0; i < got; ++i) {
        delete[] res[i].filename;
        delete[] res[i].hashprint;
    }
    delete[] res;
}

uint64_t *par_collector_calc_hashprint(LiveIdCollector *c, const char *filename, int *size) {
    if (size) *size = 0;
    try {
        if (!c || !filename || !size) throw std::invalid_argument("par_collector_calc_hashprint: NULL argument");
        auto hp = c->get().calc_hashprint(std::string(filename));
        auto *out = new uint64_t[hp.size() ? hp.size() : 1];
        std::memcpy(out, hp.data(), sizeof(uint64_t) * hp.size());
        *size = static_cast<int>(hp.size());
        return out;
    } catch (const std::exception &e) {
        hpfw_b200::set_error("par_collector_calc_hashprint: %s", e.what());
        return nullptr;
    }
}

void calc_hashprint_result_free(uint64_t *hp) { delete[] hp; }

void par_collector_save(LiveIdCollector *c, const char *cache) {
    try {
        if (!c) return;
        c->retarget(cache);
        c->get().save();
    } catch (const std::exception &e) {
        hpfw_b200::set_error("par_collector_save: %s", e.what());
    }
}

void par_collector_load(LiveIdCollector *c, const char *cache) {
    try {
        if (!c) return;
        c->retarget(cache);
        c->get().load();
    } catch (const std::exception &e) {
        hpfw_b200::set_error("par_collector_load: %s", e.what());
    }
}

}  // extern "C"
