// include/hpfw/matrix.h — minimal dense matrix with the subset of the Eigen::Matrix interface the hpfw public types use
// (rows/cols/data/resize/operator()). The reference's aliases (hashprint_handle.h:56-70, cqt.h:25) are Eigen types; Eigen is
// not a dependency of this library (the arithmetic runs on the GPU), so the aliases keep their names and memory layout:
// column-major unless RowMajor is requested.
#pragma once

#include <cstddef>
#include <vector>

namespace hpfw {

template <typename T, bool RowMajor = false>
class Matrix {
public:
    using Scalar = T;
    Matrix() = default;
    Matrix(std::ptrdiff_t rows, std::ptrdiff_t cols) { resize(rows, cols); }
    void resize(std::ptrdiff_t rows, std::ptrdiff_t cols) {
        rows_ = rows;
        cols_ = cols;
        data_.assign(static_cast<size_t>(rows) * static_cast<size_t>(cols), T());
    }
    std::ptrdiff_t rows() const { return rows_; }
    std::ptrdiff_t cols() const { return cols_; }
    std::ptrdiff_t size() const { return rows_ * cols_; }
    T *data() { return data_.data(); }
    const T *data() const { return data_.data(); }
    T &operator()(std::ptrdiff_t r, std::ptrdiff_t c) { return data_[index(r, c)]; }
    const T &operator()(std::ptrdiff_t r, std::ptrdiff_t c) const { return data_[index(r, c)]; }

private:
    size_t index(std::ptrdiff_t r, std::ptrdiff_t c) const {
        return RowMajor ? static_cast<size_t>(r) * cols_ + c : static_cast<size_t>(c) * rows_ + r;
    }
    std::ptrdiff_t rows_ = 0, cols_ = 0;
    std::vector<T> data_;
};

}  // namespace hpfw
