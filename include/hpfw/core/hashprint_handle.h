// include/hpfw/core/hashprint_handle.h — HashprintHandle: types, constants and the fused stage-2/3 entry point.
//
// Reference: /root/reference/include/hpfw/core/hashprint_handle.h:46-146. The reference exposes five static steps that
// ParallelCollector chains on the CPU (calc_frames -> filters * frames -> calc_fingerprint -> fingerprint_to_hashprint,
// parallel_collector.h:56-58). On the GPU they are ONE kernel (project.cu): the 2420 x frames context matrix and the
// 64 x frames projection are never materialised. The type aliases and constants keep the reference's names so code
// written against Algo::Spectrogram / Algo::Filters / Algo::Hashprint compiles unchanged.
#pragma once

#include <cstdint>
#include <mutex>
#include <vector>

#include "../device.h"
#include "../matrix.h"

namespace hpfw {

template <typename N, typename SpectrogramHandler, size_t FramesContext = 20, size_t T = 80, typename Real = float>
class HashprintHandle {
    static_assert(sizeof(N) == 8 && FramesContext == HPFW_CONTEXT && T == HPFW_LAG && sizeof(Real) == 4,
                  "hpfw_b200 implements HashprintHandle<uint64_t, SH, 20, 80, float> (live_song_id.h:16)");

public:
    using Spectrogram = typename SpectrogramHandler::Spectrogram;
    static constexpr size_t W = FramesContext / 2;                      // hashprint_handle.h:62
    static constexpr size_t NumOfFilters = sizeof(N) * 8;                // :64
    static constexpr size_t FrameSize = HPFW_BINS * FramesContext;       // :60
    static constexpr size_t Lag = T;

    using Frames = Matrix<Real, true>;              // row-major 2420 x frames in the reference; never built here
    using CovarianceMatrix = Matrix<Real>;          // 2420 x 2420
    using Filters = Matrix<Real>;                   // 64 x 2420, column-major, row index band*20 + context
    using Fingerprint = Matrix<bool>;               // 64 x words in the reference; never built here
    using Hashprint = std::vector<N>;

    SpectrogramHandler sh;                          // hashprint_handle.h:72

    /// Stages a4..a7 fused: spectrogram (121 x cols) -> hashprint (cols - 99 words, filter f -> bit 63-f).
    /// Throws hpfw::Error(HPFW_ERR_SHORT) for fewer than 100 columns (the reference underflows size_t there).
    static Hashprint calc_hashprint(device::Context &ctx, const Spectrogram &spectro) {
        const int cols = static_cast<int>(spectro.cols());
        const int n = hpfw_hashprint_words_for_cols(cols);
        Hashprint hp(n > 0 ? n : 0);
        int got = 0;
        std::scoped_lock l(ctx.mutex());
        device::check(hpfw_hashprint_from_spectrogram(ctx.get(), spectro.data(), cols, hp.data(), &got));
        return hp;
    }

    /// calc_filters (hashprint_handle.h:105-112): the 64 leading eigenvectors of `cov` as rows, descending eigenvalue.
    /// Sign convention: the largest-magnitude component of every filter is positive (solvers differ in sign anyway).
    static Filters calc_filters(device::Context &ctx, const CovarianceMatrix &cov) {
        if (cov.rows() != static_cast<std::ptrdiff_t>(FrameSize) || cov.cols() != static_cast<std::ptrdiff_t>(FrameSize))
            throw Error(HPFW_ERR_ARG, "covariance must be 2420 x 2420");
        Filters f(NumOfFilters, FrameSize);
        std::scoped_lock l(ctx.mutex());
        device::check(hpfw_calc_filters(ctx.get(), cov.data(), f.data(), nullptr));
        return f;
    }

    /// Upload the learned filters (64 x 2420 column-major, as Filters stores them) to the device context.
    static void set_filters(device::Context &ctx, const Filters &filters) {
        if (filters.rows() != static_cast<std::ptrdiff_t>(NumOfFilters) ||
            filters.cols() != static_cast<std::ptrdiff_t>(FrameSize))
            throw Error(HPFW_ERR_ARG, "filters must be 64 x 2420");
        std::scoped_lock l(ctx.mutex());
        device::check(hpfw_set_filters(ctx.get(), filters.data()));
    }
};

}  // namespace hpfw
