// examples/cpp/live-id.cpp — the reference's example (examples/cpp/live-id.cpp:12-25, README.md:12-33) against the B200 path.
//
//   live-id <index_dir> <search_dir> [filters.cereal]
//
// index_dir / search_dir hold mono or stereo WAV files at HPFW_EXAMPLE_RATE Hz (default 44100; BASELINE config 0 uses
// -DHPFW_EXAMPLE_RATE=22050). Filters come from cache/filters.cereal (written by
// hpfw itself, or copied there by the optional third argument).
//
// Build:  g++ -std=c++17 -O2 -Iinclude examples/cpp/live-id.cpp -o live-id -Lhpfw_b200 -lhpfw_b200 -Wl,-rpath,$PWD/hpfw_b200
#include <algorithm>
#include <filesystem>
#include <iostream>

#include <hpfw/audioproblems/live-song-id/live_song_id.h>

int main(int argc, char **argv) {
    if (argc < 3) {
        std::cerr << "usage: " << argv[0] << " <index_dir> <search_dir> [filters.cereal]" << std::endl;
        return 2;
    }
    std::ios_base::sync_with_stdio(false);
    std::cin.tie(nullptr);
    try {
        if (argc > 3) {
            std::filesystem::create_directories("cache/spectros");
            std::filesystem::copy_file(argv[3], "cache/filters.cereal", std::filesystem::copy_options::overwrite_existing);
        }
        auto index_files = hpfw::utils::get_dir_files(argv[1]);
        auto search_files = hpfw::utils::get_dir_files(argv[2]);
        std::sort(search_files.begin(), search_files.end());

#ifdef HPFW_EXAMPLE_RATE
        using Algo = hpfw::HashprintHandle<uint64_t, hpfw::spectrum::CQT<HPFW_EXAMPLE_RATE>, 20, 80>;
        using Coll = hpfw::ParallelCollector<Algo, hpfw::cache::DriveCache>;
        hpfw::LiveSongIdentification<Coll> liveid;
#else
        hpfw::LiveSongIdentification liveid;
#endif
        liveid.index(index_files);
        auto s = liveid.search(search_files);
        return s.failed == 0 ? 0 : 1;
    } catch (const std::exception &e) {
        std::cerr << "live-id: " << e.what() << std::endl;
        return 1;
    }
}
