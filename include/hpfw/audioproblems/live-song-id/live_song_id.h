// include/hpfw/audioproblems/live-song-id/live_song_id.h — hpfw::LiveSongIdentification, the public API of the path.
//
// Same class, defaults-by-alias and index()/search() as /root/reference/include/hpfw/audioproblems/live-song-id/
// live_song_id.h:16-60, with db::MemoryStorage as the default Storage (the reference's default names db::AnnStorage through
// a header that does not exist, live_song_id.h:12; north_star specifies the exhaustive XOR+popcount matcher).
#pragma once

#include <filesystem>
#include <iostream>
#include <string>
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

    /// Serial query loop with the reference's printout and file-name accuracy heuristic (:35-54).
    SearchSummary search(const std::vector<std::string> &filenames) {
        SearchSummary s;
        s.queries = filenames.size();
        for (const auto &f : filenames) {
            std::cout << "=> Finding " << f << std::endl;
            try {
                auto res = storage.find(collector.calc_hashprint(f));
                auto res_name = std::filesystem::path(res.filename).stem().string();
                if (f.find(res_name) == std::string::npos) {
                    std::cerr << "[hpfw] Wrong result for '" << f << "': got '" << res_name << "'" << std::endl;
                    ++s.wrong;
                }
                std::cout << "=> " << res.filename << " " << res.cnt << " " << res.offset << std::endl << std::endl;
            } catch (const std::exception &e) {
                std::cerr << "[hpfw] Error finding '" << f << "': " << e.what() << std::endl;
                ++s.failed;
            }
        }
        std::cout << "=> " << s.wrong << " " << s.accuracy() << std::endl;
        return s;
    }

    Collector &get_collector() { return collector; }
    Storage &get_storage() { return storage; }

private:
    Collector collector;
    Storage storage;
};

}  // namespace hpfw
