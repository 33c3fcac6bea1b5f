// include/hpfw/io/cereal_compat.h — readers/writers for the reference's on-disk files WITHOUT cereal.
// cereal's BinaryOutputArchive writes raw little-endian bytes with no header, so the layouts are:
//   Eigen matrix (utils.h:77-106):  int32 rows, int32 cols, rows*cols scalars in the matrix's storage order
//       cache/filters.cereal   = 64 x 2420 float, column-major      (cache.h:39-41, hashprint_handle.h:68)
//       cache/accum_cov.cereal = 2420 x 2420 float, column-major    (cache.h:35-37)
//       cache/spectros/<stem>  = 121 x cols float, column-major     (cache.h:30-33, cqt.h:25)
//   MemoryStorage dump (storage.h:67-86) = std::vector<FilenameFingerprintPair>:
//       uint64 n; n x { uint64 len, len bytes of filename; uint64 cnt, cnt x uint64 hashprint words }
#pragma once

#include <cstdint>
#include <filesystem>
#include <fstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../matrix.h"

namespace hpfw::io {

template <typename T, bool RM>
inline void save_matrix(const std::string &filename, const Matrix<T, RM> &m) {
    std::ofstream os(filename, std::ios::binary);
    if (!os) throw std::runtime_error("cannot write '" + filename + "'");
    const int32_t rows = static_cast<int32_t>(m.rows()), cols = static_cast<int32_t>(m.cols());
    os.write(reinterpret_cast<const char *>(&rows), 4);
    os.write(reinterpret_cast<const char *>(&cols), 4);
    os.write(reinterpret_cast<const char *>(m.data()), static_cast<std::streamsize>(sizeof(T)) * rows * cols);
}

/// save_matrix for a raw buffer (float, column-major rows x cols); written to a temporary name and renamed, so that a reader
/// never sees a half-written cache file.
inline void save_matrix_raw(const std::string &filename, const float *data, int32_t rows, int32_t cols) {
    const std::string tmp = filename + ".tmp~";
    {
        std::ofstream os(tmp, std::ios::binary);
        if (!os) throw std::runtime_error("cannot write '" + tmp + "'");
        os.write(reinterpret_cast<const char *>(&rows), 4);
        os.write(reinterpret_cast<const char *>(&cols), 4);
        os.write(reinterpret_cast<const char *>(data), static_cast<std::streamsize>(sizeof(float)) * rows * cols);
        if (!os) throw std::runtime_error("cannot write '" + tmp + "'");
    }
    std::filesystem::rename(tmp, filename);
}

/// Header of a matrix file: (rows, cols); the payload follows at byte 8.
inline bool matrix_header(const std::string &filename, int32_t &rows, int32_t &cols) {
    std::ifstream is(filename, std::ios::binary);
    if (!is) return false;
    is.read(reinterpret_cast<char *>(&rows), 4);
    is.read(reinterpret_cast<char *>(&cols), 4);
    return bool(is) && rows >= 0 && cols >= 0;
}

/// Payload of a matrix file straight into `dst` (rows * cols scalars of sizeof(T)).
template <typename T>
inline void load_matrix_payload(const std::string &filename, T *dst, size_t count) {
    std::ifstream is(filename, std::ios::binary);
    is.seekg(8);
    is.read(reinterpret_cast<char *>(dst), static_cast<std::streamsize>(sizeof(T) * count));
    if (!is) throw std::runtime_error("'" + filename + "': truncated matrix file");
}

/// Like DriveCache::load (cache.h:73-82): a missing file leaves `m` untouched and returns false.
template <typename T, bool RM>
inline bool load_matrix(const std::string &filename, Matrix<T, RM> &m) {
    if (!std::filesystem::exists(filename)) return false;
    std::ifstream is(filename, std::ios::binary);
    int32_t rows = 0, cols = 0;
    is.read(reinterpret_cast<char *>(&rows), 4);
    is.read(reinterpret_cast<char *>(&cols), 4);
    if (!is || rows < 0 || cols < 0) throw std::runtime_error("'" + filename + "': bad matrix header");
    m.resize(rows, cols);
    is.read(reinterpret_cast<char *>(m.data()), static_cast<std::streamsize>(sizeof(T)) * rows * cols);
    if (!is) throw std::runtime_error("'" + filename + "': truncated matrix file");
    return true;
}

using NamedHashprint = std::pair<std::string, std::vector<uint64_t>>;

inline void save_db(const std::string &filename, const std::vector<NamedHashprint> &db) {
    std::ofstream os(filename, std::ios::binary);
    if (!os) throw std::runtime_error("cannot write '" + filename + "'");
    const uint64_t n = db.size();
    os.write(reinterpret_cast<const char *>(&n), 8);
    for (const auto &e : db) {
        const uint64_t len = e.first.size(), cnt = e.second.size();
        os.write(reinterpret_cast<const char *>(&len), 8);
        os.write(e.first.data(), static_cast<std::streamsize>(len));
        os.write(reinterpret_cast<const char *>(&cnt), 8);
        os.write(reinterpret_cast<const char *>(e.second.data()), static_cast<std::streamsize>(cnt * 8));
    }
}

inline std::vector<NamedHashprint> load_db(const std::string &filename) {
    std::ifstream is(filename, std::ios::binary);
    if (!is) throw std::runtime_error("cannot open '" + filename + "'");
    uint64_t n = 0;
    is.read(reinterpret_cast<char *>(&n), 8);
    std::vector<NamedHashprint> db;
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t len = 0, cnt = 0;
        is.read(reinterpret_cast<char *>(&len), 8);
        if (!is || len > (1u << 20)) throw std::runtime_error("'" + filename + "': bad DB dump");
        std::string name(len, '\0');
        is.read(name.data(), static_cast<std::streamsize>(len));
        is.read(reinterpret_cast<char *>(&cnt), 8);
        if (!is || cnt > (uint64_t(1) << 32)) throw std::runtime_error("'" + filename + "': bad DB dump");
        std::vector<uint64_t> hp(cnt);
        is.read(reinterpret_cast<char *>(hp.data()), static_cast<std::streamsize>(cnt * 8));
        if (!is) throw std::runtime_error("'" + filename + "': truncated DB dump");
        db.emplace_back(std::move(name), std::move(hp));
    }
    return db;
}

}  // namespace hpfw::io
