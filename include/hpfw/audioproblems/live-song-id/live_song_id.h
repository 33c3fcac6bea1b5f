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

    /// index (:31-33). When the collector can leave the hashprints in HBM (prepare_device) and the storage can take them from
    /// there (build_device), the database is built device-to-device; any other Collector / Storage plug-in gets the
    /// reference's storage.build(collector.prepare(files)).
    void index(const std::vector<std::string> &filenames) { index_impl(collector, storage, filenames, 0); }

    /// The reference's query loop, printout and file-name accuracy heuristic (:35-54). The reference matches one query at
    /// a time; here the hashprints of all query files are extracted as one batch (N decode threads -> pinned ring -> GPU,
    /// left in HBM when collector and storage support it) and matched in ONE batched call when the storage offers find_topk
    /// (db::MemoryStorage does: groups of 128 queries run as an exact GEMM on the tensor cores, match_tc.cu). Results, their
    /// order and the text written to stdout are the same as with the serial loop. A file that cannot be read, is too short or
    /// exceeds the matcher's query length fails alone, as with the reference's per-query try/catch (:40-51).
    SearchSummary search(const std::vector<std::string> &filenames) {
        SearchSummary s;
        s.queries = filenames.size();
        using Hashprint = decltype(collector.calc_hashprint(std::declval<const std::string &>()));
        using Result = decltype(storage.find(std::declval<const Hashprint &>()));
        std::vector<Result> results(filenames.size());
        std::vector<std::string> error(filenames.size());
        search_impl(collector, storage, filenames, results, error, 0);
        for (size_t i = 0; i < filenames.size(); ++i) {
            const auto &f = filenames[i];
            std::cout << "=> Finding " << f << std::endl;
            if (!error[i].empty()) {
                std::cerr << "[hpfw] Error finding '" << f << "': " << error[i] << std::endl;
                ++s.failed;
                continue;
            }
            const auto &res = results[i];
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
    // ---- index: device-to-device when both plug-ins support it
    template <class C, class St>
    static auto index_impl(C &c, St &st, const std::vector<std::string> &files, int)
        -> decltype(st.build_device(c.prepare_device(files, false)), c.begin_cache_writes(), void()) {
        detail::PhaseTrace trace;
        auto d = c.prepare_device(files, false);       // the collector's cache writers wait ...
        trace.mark("index: prepare_device");
        st.build_device(d);
        trace.mark("index: storage build_device");
        c.begin_cache_writes();                        // ... until the storage is built
    }
    template <class C, class St>
    static void index_impl(C &c, St &st, const std::vector<std::string> &files, long) {
        st.build(c.prepare(files));
    }

    // ---- search, best path first: (1) hashprints stay in HBM between extraction and match
    template <class C, class St, class R>
    static auto search_impl(C &c, St &st, const std::vector<std::string> &files, std::vector<R> &results,
                            std::vector<std::string> &error, int)
        -> decltype(st.find_topk_device(c.calc_hashprints_device(files), 1), void()) {
        auto d = c.calc_hashprints_device(files);
        error = d.errors;
        std::string batch_error;
        decltype(st.find_topk_device(d, 1)) found;
        try {
            found = st.find_topk_device(d, 1);
        } catch (const std::exception &e) {
            batch_error = e.what();
        }
        for (size_t i = 0; i < files.size(); ++i) {
            if (!error[i].empty()) continue;
            if (!batch_error.empty() || d.slot_of_input[i] < 0) error[i] = batch_error.empty() ? "not extracted" : batch_error;
            else results[i] = std::move(found[static_cast<size_t>(d.slot_of_input[i])].at(0));
        }
    }
    // (2) any collector: one calc_hashprint per file, then one batched find_topk if the storage has it, else find per query
    template <class C, class St, class R>
    static void search_impl(C &c, St &st, const std::vector<std::string> &files, std::vector<R> &results,
                            std::vector<std::string> &error, long) {
        using Hashprint = decltype(c.calc_hashprint(std::declval<const std::string &>()));
        std::vector<Hashprint> hps;
        std::vector<size_t> input_of;
        for (size_t i = 0; i < files.size(); ++i) {
            try {
                hps.push_back(c.calc_hashprint(files[i]));
                input_of.push_back(i);
            } catch (const std::exception &e) {
                error[i] = e.what();
                if (error[i].empty()) error[i] = "calc_hashprint failed";
            }
        }
        bool batched_ok = false;
        try {
            batched_ok = find_batch(st, hps, input_of, results, 0);
        } catch (const std::exception &) {
            batched_ok = false;     // one bad query must not fail the others: fall through to one find() per query
        }
        if (!batched_ok) {
            for (size_t k = 0; k < hps.size(); ++k) {
                try {
                    results[input_of[k]] = st.find(hps[k]);
                } catch (const std::exception &e) {
                    error[input_of[k]] = e.what();
                    if (error[input_of[k]].empty()) error[input_of[k]] = "find failed";
                }
            }
        }
    }
    template <class St, class Hp, class R>
    static auto find_batch(St &st, const std::vector<Hp> &hps, const std::vector<size_t> &input_of, std::vector<R> &results, int)
        -> decltype(st.find_topk(hps, 1), bool()) {
        if (hps.empty()) return true;
        auto found = st.find_topk(hps, 1);
        for (size_t k = 0; k < hps.size(); ++k) results[input_of[k]] = std::move(found[k].at(0));
        return true;
    }
    template <class St, class Hp, class R>
    static bool find_batch(St &, const std::vector<Hp> &, const std::vector<size_t> &, std::vector<R> &, long) {
        return false;
    }

    Collector collector;
    Storage storage;
};

}  // namespace hpfw
