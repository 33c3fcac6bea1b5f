// include/hpfw/utils.h — host helpers the README usage relies on (reference: include/hpfw/utils.h:36-55).
#pragma once

#include <algorithm>
#include <cstdint>
#include <filesystem>
#include <string>
#include <vector>

namespace hpfw::utils {

/// All directory entries of `dir` as paths (reference utils.h:36-44). Sorted, so DB order is reproducible (the reference
/// leaves it to directory_iterator order).
inline std::vector<std::string> get_dir_files(const std::string &dir) {
    std::vector<std::string> files;
    for (const auto &f : std::filesystem::directory_iterator(dir)) files.emplace_back(f.path().string());
    std::sort(files.begin(), files.end());
    return files;
}

/// Number of regular files in `dir` (reference utils.h:46-55).
inline uint64_t count_dir_files(const std::string &dir) {
    uint64_t n = 0;
    for (const auto &f : std::filesystem::directory_iterator(dir))
        if (f.is_regular_file()) ++n;
    return n;
}

}  // namespace hpfw::utils
