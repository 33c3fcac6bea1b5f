// include/hpfw/audioproblems/live-song-id/live_song_id.h — hpfw::LiveSongIdentification, the public API of the path.
//
// Same class, defaults-by-alias and index()/search() as /root/reference/include/hpfw/audioproblems/live-song-id/
// live_song_id.h:16-60, with db::MemoryStorage as the default Storage (the reference's default names db::AnnStorage through
// a header that does not exist, live_song_id.h:12; north_star specifies the exhaustive XOR+popcount matcher).
#pragma once

#include <cstdint>
#include <filesystem>
#include <iostream>
#include <string>
#include <utility>
#include <vector>

#include "../../core/cache.h"
#include "../../core/hashprint_handle.h"
#include "../../core/parallel_collector.h"
#include "../../spectrum/cqt.h"
#include "storage.h"

namespace hpfw {

using DefaultLiveIdAlgoConfig = HashprintHandle<uint64_t, spectrum::CQT<>, 20, 80>;
using DefaultLiveIdCollector = ParallelCollector<DefaultLiveIdAlgoConfig, cache::DriveCache>;

template <typename Collector = DefaultLiveIdCollector, typename Storage = db::MemoryStorage<Collector>>
class LiveSongIdentification {
public:
    struct SearchSummary {
        size_t queries = 0, wrong = 0, failed = 0;
        float accuracy() const { return queries ? 1.f - wrong / float(queries) : 0.f; }
    };

    LiveSongIdentification() { collector.load(); }      // live_song_id.h:23-25
    ~LiveSongIdentification() {
        try { collector.save(); } catch (...) {}         // :27-29
    }

    void index(const std::vector<std::string> &filenames) { storage.build(collector.prepare(filenames)); }   // :31-33

    /// The reference's query loop, printout and file-name accuracy heuristic (:35-54). The reference matches one query at
    /// a time; here the hashprints of all query files are computed first and matched in ONE batched call when the storage
    /// offers find_topk (db::MemoryStorage does: groups of 128 queries run as an exact GEMM on the tensor cores, match_tc.cu).
    /// Results, their order and the text written to stdout are the same as with the serial loop.
    SearchSummary search(const std::vector<std::string> &filenames) {
        SearchSummary s;
        s.queries = filenames.size();
        using Hashprint = decltype(collector.calc_hashprint(std::declval<const std::string &>()));
        using Result = decltype(storage.find(std::declval<const Hashprint &>()));
        std::vector<Hashprint> hps;
        std::vector<size_t> slot(filenames.size(), SIZE_MAX);
        std::vector<std::string> error(filenames.size());
        for (size_t i = 0; i < filenames.size(); ++i) {
            try {
                hps.push_back(collector.calc_hashprint(filenames[i]));
                slot[i] = hps.size() - 1;
            } catch (const std::exception &e) {
                error[i] = e.what();
            }
        }
        std::vector<Result> results;
        std::string batch_error;
        try {
            results = find_all(storage, hps, 0);
        } catch (const std::exception &e) {
            batch_error = e.what();
        }
        for (size_t i = 0; i < filenames.size(); ++i) {
            const auto &f = filenames[i];
            std::cout << "=> Finding " << f << std::endl;
            if (slot[i] == SIZE_MAX || !batch_error.empty()) {
                std::cerr << "[hpfw] Error finding '" << f << "': " << (slot[i] == SIZE_MAX ? error[i] : batch_error)
                          << std::endl;
                ++s.failed;
                continue;
            }
            const auto &res = results[slot[i]];
            auto res_name = std::filesystem::path(res.filename).stem().string();
            if (f.find(res_name) == std::string::npos) {
                std::cerr << "[hpfw] Wrong result for '" << f << "': got '" << res_name << "'" << std::endl;
                ++s.wrong;
            }
            std::cout << "=> " << res.filename << " " << res.cnt << " " << res.offset << std::endl << std::endl;
        }
        std::cout << "=> " << s.wrong << " " << s.accuracy() << std::endl;
        return s;
    }

    Collector &get_collector() { return collector; }
    Storage &get_storage() { return storage; }

private:
    // one batched call when the storage plug-in has find_topk(queries, k), else the reference's one find() per query
    template <class St, class Hp>
    static auto find_all(St &st, const std::vector<Hp> &hps, int) -> decltype(st.find_topk(hps, 1), std::vector<decltype(st.find(hps[0]))>()) {
        std::vector<decltype(st.find(hps[0]))> out;
        if (hps.empty()) return out;
        for (auto &r : st.find_topk(hps, 1)) out.push_back(std::move(r.at(0)));
        return out;
    }
    template <class St, class Hp>
    static auto find_all(St &st, const std::vector<Hp> &hps, long) -> std::vector<decltype(st.find(hps[0]))> {
        std::vector<decltype(st.find(hps[0]))> out;
        for (const auto &hp : hps) out.push_back(st.find(hp));
        return out;
    }

    Collector collector;
    Storage storage;
};

}  // namespace hpfw
