// include/hpfw/io/wav.h — RIFF/WAVE reader for the file-name entry points (spectrum::CQT::spectrogram(filename)).
// The reference decodes any format through essentia's MonoLoader (ffmpeg) and resamples to SampleRate (cqt.h:45-52);
// there is no decoder library here, so: PCM 8/16/24/32-bit or IEEE float32/64 WAV, any channel count (down-mixed to mono by
// averaging, as MonoLoader's default "mix"), and the file's rate must already equal the requested rate (no resampler).
#pragma once

#include <cstdint>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace hpfw::io {

struct WavData {
    int sample_rate = 0;
    std::vector<float> mono;
    std::vector<int16_t> pcm16;   // filled INSTEAD of `mono` when keep_pcm16 was asked for and the file is mono 16-bit PCM
};

/// keep_pcm16: a mono 16-bit PCM file is returned as its raw samples (`pcm16`; one bulk read, no per-sample work), for the
/// hpfw_*_pcm16 entry points that convert on the device. Everything else is converted to float here.
inline WavData read_wav(const std::string &filename, bool keep_pcm16 = false) {
    std::ifstream is(filename, std::ios::binary);
    if (!is) throw std::runtime_error("cannot open '" + filename + "'");
    auto rd = [&](void *p, size_t n) {
        is.read(static_cast<char *>(p), static_cast<std::streamsize>(n));
        if (static_cast<size_t>(is.gcount()) != n) throw std::runtime_error("'" + filename + "': truncated WAV file");
    };
    char riff[12];
    rd(riff, 12);
    if (std::memcmp(riff, "RIFF", 4) != 0 || std::memcmp(riff + 8, "WAVE", 4) != 0)
        throw std::runtime_error("'" + filename + "': not a RIFF/WAVE file (only WAV input is supported)");
    uint16_t fmt = 0, channels = 0, bits = 0;
    uint32_t rate = 0;
    bool have_fmt = false;
    WavData out;
    for (;;) {
        char id[4];
        uint32_t sz = 0;
        is.read(id, 4);
        if (is.gcount() != 4) break;
        rd(&sz, 4);
        if (std::memcmp(id, "fmt ", 4) == 0) {
            std::vector<unsigned char> f(sz);
            rd(f.data(), sz);
            if (sz < 16) throw std::runtime_error("'" + filename + "': bad fmt chunk");
            std::memcpy(&fmt, &f[0], 2);
            std::memcpy(&channels, &f[2], 2);
            std::memcpy(&rate, &f[4], 4);
            std::memcpy(&bits, &f[14], 2);
            if (fmt == 0xFFFE && sz >= 26) std::memcpy(&fmt, &f[24], 2);   // WAVE_FORMAT_EXTENSIBLE: sub-format
            have_fmt = true;
        } else if (std::memcmp(id, "data", 4) == 0) {
            if (!have_fmt || channels == 0) throw std::runtime_error("'" + filename + "': data before fmt");
            std::vector<unsigned char> raw(sz);
            is.read(reinterpret_cast<char *>(raw.data()), sz);
            raw.resize(static_cast<size_t>(is.gcount()));
            const size_t bps = bits / 8, frame = bps * channels;
            if (bps == 0) throw std::runtime_error("'" + filename + "': bad bit depth");
            const size_t n = raw.size() / frame;
            out.sample_rate = static_cast<int>(rate);
            if (channels == 1 && fmt == 1 && bits == 16 && keep_pcm16) {
                out.pcm16.resize(n);
                std::memcpy(out.pcm16.data(), raw.data(), n * 2);
                return out;
            }
            out.mono.resize(n);
            if (channels == 1 && fmt == 3 && bits == 32) {        // mono float32: the samples as they are
                std::memcpy(out.mono.data(), raw.data(), n * 4);
                return out;
            }
            if (channels == 1 && fmt == 1 && bits == 16) {        // mono 16-bit PCM: MonoLoader's sample / 32768, exact in float
                for (size_t i = 0; i < n; ++i) {
                    int16_t v;
                    std::memcpy(&v, &raw[2 * i], 2);
                    out.mono[i] = static_cast<float>(v) * (1.0f / 32768.0f);
                }
                return out;
            }
            for (size_t i = 0; i < n; ++i) {
                double acc = 0.0;
                for (unsigned c = 0; c < channels; ++c) {
                    const unsigned char *p = &raw[i * frame + c * bps];
                    double v = 0.0;
                    if (fmt == 3 && bits == 32) { float f; std::memcpy(&f, p, 4); v = f; }
                    else if (fmt == 3 && bits == 64) { double d; std::memcpy(&d, p, 8); v = d; }
                    else if (fmt == 1 && bits == 8) v = (static_cast<int>(p[0]) - 128) / 128.0;
                    else if (fmt == 1 && bits == 16) { int16_t s; std::memcpy(&s, p, 2); v = s / 32768.0; }
                    else if (fmt == 1 && bits == 24) {
                        int32_t s = (p[0] << 8) | (p[1] << 16) | (static_cast<int32_t>(p[2]) << 24);
                        v = (s >> 8) / 8388608.0;
                    } else if (fmt == 1 && bits == 32) { int32_t s; std::memcpy(&s, p, 4); v = s / 2147483648.0; }
                    else throw std::runtime_error("'" + filename + "': unsupported WAV sample format");
                    acc += v;
                }
                out.mono[i] = static_cast<float>(acc / channels);
            }
            out.sample_rate = static_cast<int>(rate);
            return out;
        } else {
            is.seekg(sz + (sz & 1), std::ios::cur);
        }
    }
    throw std::runtime_error("'" + filename + "': no data chunk");
}

/// Writes mono float32 WAV (used by the examples and tests to materialise synthetic audio).
inline void write_wav_f32(const std::string &filename, const float *samples, size_t n, int sample_rate) {
    std::ofstream os(filename, std::ios::binary);
    if (!os) throw std::runtime_error("cannot create '" + filename + "'");
    const uint32_t data_bytes = static_cast<uint32_t>(n * 4), riff = 36 + data_bytes, fmt_sz = 16, rate = sample_rate,
                   byte_rate = rate * 4;
    const uint16_t fmt = 3, ch = 1, align = 4, bits = 32;
    os.write("RIFF", 4); os.write(reinterpret_cast<const char *>(&riff), 4); os.write("WAVEfmt ", 8);
    os.write(reinterpret_cast<const char *>(&fmt_sz), 4); os.write(reinterpret_cast<const char *>(&fmt), 2);
    os.write(reinterpret_cast<const char *>(&ch), 2); os.write(reinterpret_cast<const char *>(&rate), 4);
    os.write(reinterpret_cast<const char *>(&byte_rate), 4); os.write(reinterpret_cast<const char *>(&align), 2);
    os.write(reinterpret_cast<const char *>(&bits), 2); os.write("data", 4);
    os.write(reinterpret_cast<const char *>(&data_bytes), 4);
    os.write(reinterpret_cast<const char *>(samples), data_bytes);
}

}  // namespace hpfw::io
