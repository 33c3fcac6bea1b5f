// include/hpfw/spectrum/cqt.h — spectrum::CQT, the SpectrogramHandler plug-in of HashprintHandle.
// Same name, template parameters, Spectrogram alias and static spectrogram(filename) as the reference
// (/root/reference/include/hpfw/spectrum/cqt.h:18-36); the transform itself runs on the GPU (hpfw_cqt_spectrogram,
// hpfw_b200/csrc/cqt.cu) instead of essentia/FFTW.
#pragma once

#include <cstdint>
#include <string>
#include <utility>

#include "../device.h"
#include "../io/wav.h"
#include "../matrix.h"

namespace hpfw::spectrum {

template <uint32_t SampleRate = 44100, uint32_t HopLength = 96, uint32_t BinsPerOctave = 24, uint32_t NumberBins = 121,
          uint32_t DownsampleFactor = 3>
class CQT {
    // The kernels implement the reference's only instantiation of the transform geometry (live_song_id.h:16); SampleRate
    // is free because the reference never forwards it to NSGConstantQ (cqt.h:54-61) — it only selects the decode rate.
    static_assert(HopLength == 96 && BinsPerOctave == 24 && NumberBins == 121 && DownsampleFactor == 3,
                  "hpfw_b200 implements CQT<SR, 96, 24, 121, 3> (the reference's default geometry)");

public:
    using Spectrogram = Matrix<float>;   // 121 x cols, column-major (reference: Eigen::Matrix<float, 121, Dynamic>)
    static constexpr uint32_t sample_rate = SampleRate;

    CQT() = default;

    /// Reference signature (cqt.h:36). Decodes a WAV file whose rate must equal SampleRate (no resampler here).
    static Spectrogram spectrogram(const std::string &filename, int device = 0) {
        const io::WavData wav = io::read_wav(filename, /*keep_pcm16=*/true);
        if (wav.sample_rate != static_cast<int>(SampleRate))
            throw Error(HPFW_ERR_ARG, "'" + filename + "' is sampled at " + std::to_string(wav.sample_rate) +
                                          " Hz; CQT<" + std::to_string(SampleRate) + "> needs that rate (no resampler)");
        if (!wav.pcm16.empty()) return spectrogram(wav.pcm16.data(), static_cast<int64_t>(wav.pcm16.size()), device);
        return spectrogram(wav.mono.data(), static_cast<int64_t>(wav.mono.size()), device);
    }

    /// Decode hook of the batched path (ParallelCollector::prepare / LiveSongIdentification::search): the samples go straight
    /// into memory provided by `alloc(bytes)` — a pinned staging slot of the extraction stream — as mono int16 or float32.
    template <typename Alloc>
    static io::WavInfo decode(const std::string &filename, Alloc &&alloc) {
        const io::WavInfo info = io::read_wav_into(filename, std::forward<Alloc>(alloc));
        if (info.sample_rate != static_cast<int>(SampleRate))
            throw Error(HPFW_ERR_ARG, "'" + filename + "' is sampled at " + std::to_string(info.sample_rate) +
                                          " Hz; CQT<" + std::to_string(SampleRate) + "> needs that rate (no resampler)");
        return info;
    }

    /// Mono 16-bit PCM samples: copied as they are (half the bytes), MonoLoader's sample / 32768 is done on the device.
    static Spectrogram spectrogram(const int16_t *pcm, int64_t n_samples, int device = 0) {
        auto ctx = device::Context::shared(device);
        std::scoped_lock l(ctx->mutex());
        const int cols = hpfw_cqt_cols(n_samples);
        Spectrogram s(NumberBins, cols > 0 ? cols : 0);
        int got = 0;
        device::check(hpfw_cqt_spectrogram_pcm16(ctx->get(), pcm, n_samples, s.data(), &got));
        return s;
    }

    /// Same on an already decoded mono buffer.
    static Spectrogram spectrogram(const float *audio, int64_t n_samples, int device = 0) {
        auto ctx = device::Context::shared(device);
        std::scoped_lock l(ctx->mutex());
        const int cols = hpfw_cqt_cols(n_samples);
        Spectrogram s(NumberBins, cols > 0 ? cols : 0);
        int got = 0;
        device::check(hpfw_cqt_spectrogram(ctx->get(), audio, n_samples, s.data(), &got));
        return s;
    }
};

}  // namespace hpfw::spectrum
