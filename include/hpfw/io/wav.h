// include/hpfw/io/wav.h — RIFF/WAVE reader for the file-name entry points (spectrum::CQT::spectrogram(filename)).
// The reference decodes any format through essentia's MonoLoader (ffmpeg) and resamples to SampleRate (cqt.h:45-52);
// there is no decoder library here, so: PCM 8/16/24/32-bit or IEEE float32/64 WAV, any channel count (down-mixed to mono by
// averaging, as MonoLoader's default "mix"), and the file's rate must already equal the requested rate (no resampler).
#pragma once

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace hpfw::io {

namespace detail {
/// one sample of any supported WAV encoding as a double in [-1, 1)
inline double wav_sample(const unsigned char *p, uint16_t fmt, uint16_t bits, const std::string &filename) {
    if (fmt == 3 && bits == 32) { float f; std::memcpy(&f, p, 4); return f; }
    if (fmt == 3 && bits == 64) { double d; std::memcpy(&d, p, 8); return d; }
    if (fmt == 1 && bits == 8) return (static_cast<int>(p[0]) - 128) / 128.0;
    if (fmt == 1 && bits == 16) { int16_t s; std::memcpy(&s, p, 2); return s / 32768.0; }
    if (fmt == 1 && bits == 24) {
        const int32_t s = (p[0] << 8) | (p[1] << 16) | (static_cast<int32_t>(p[2]) << 24);
        return (s >> 8) / 8388608.0;
    }
    if (fmt == 1 && bits == 32) { int32_t s; std::memcpy(&s, p, 4); return s / 2147483648.0; }
    throw std::runtime_error("'" + filename + "': unsupported WAV sample format");
}
}  // namespace detail

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
                    acc += detail::wav_sample(&raw[i * frame + c * bps], fmt, bits, filename);
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

struct WavInfo {
    int sample_rate = 0;
    int64_t n_samples = 0;
    bool pcm16 = false;   // the destination holds int16 samples (mono 16-bit PCM file); else mono float32
};

/// Decode straight into caller-provided memory: `alloc(bytes)` is called once and must return a buffer of that size (the
/// extraction stream hands out a pinned staging slot, so a mono 16-bit PCM or float32 file goes disk -> pinned memory in one
/// read(2), with no intermediate copy and no per-sample work on the host). Other layouts are converted and down-mixed to
/// float as read_wav does.
template <typename Alloc>
inline WavInfo read_wav_into(const std::string &filename, Alloc &&alloc) {
    struct Fd {
        int fd;
        ~Fd() { if (fd >= 0) ::close(fd); }
    } f{::open(filename.c_str(), O_RDONLY | O_CLOEXEC)};
    if (f.fd < 0) throw std::runtime_error("cannot open '" + filename + "'");
    off_t pos = 0;
    auto rd = [&](void *p, size_t n, bool exact) -> size_t {
        size_t got = 0;
        while (got < n) {
            const ssize_t r = ::pread(f.fd, static_cast<char *>(p) + got, n - got, pos + static_cast<off_t>(got));
            if (r < 0) throw std::runtime_error("'" + filename + "': read error");
            if (r == 0) break;
            got += static_cast<size_t>(r);
        }
        if (exact && got != n) throw std::runtime_error("'" + filename + "': truncated WAV file");
        pos += static_cast<off_t>(got);
        return got;
    };
    char riff[12];
    rd(riff, 12, true);
    if (std::memcmp(riff, "RIFF", 4) != 0 || std::memcmp(riff + 8, "WAVE", 4) != 0)
        throw std::runtime_error("'" + filename + "': not a RIFF/WAVE file (only WAV input is supported)");
    uint16_t fmt = 0, channels = 0, bits = 0;
    uint32_t rate = 0;
    bool have_fmt = false;
    for (;;) {
        char hdr[8];
        if (rd(hdr, 8, false) != 8) break;
        uint32_t sz = 0;
        std::memcpy(&sz, hdr + 4, 4);
        if (std::memcmp(hdr, "fmt ", 4) == 0) {
            unsigned char fb[40] = {};
            if (sz < 16) throw std::runtime_error("'" + filename + "': bad fmt chunk");
            const size_t take = sz < sizeof(fb) ? sz : sizeof(fb);
            rd(fb, take, true);
            pos += static_cast<off_t>(sz - take) + (sz & 1);
            std::memcpy(&fmt, &fb[0], 2);
            std::memcpy(&channels, &fb[2], 2);
            std::memcpy(&rate, &fb[4], 4);
            std::memcpy(&bits, &fb[14], 2);
            if (fmt == 0xFFFE && sz >= 26) std::memcpy(&fmt, &fb[24], 2);   // WAVE_FORMAT_EXTENSIBLE: sub-format
            have_fmt = true;
        } else if (std::memcmp(hdr, "data", 4) == 0) {
            if (!have_fmt || channels == 0) throw std::runtime_error("'" + filename + "': data before fmt");
            const size_t bps = bits / 8, frame = bps * channels;
            if (bps == 0) throw std::runtime_error("'" + filename + "': bad bit depth");
            struct stat stt;
            size_t avail = sz;
            if (::fstat(f.fd, &stt) == 0 && stt.st_size > pos) avail = std::min<size_t>(sz, static_cast<size_t>(stt.st_size - pos));
            const size_t n = avail / frame;
            WavInfo info;
            info.sample_rate = static_cast<int>(rate);
            info.n_samples = static_cast<int64_t>(n);
            if (channels == 1 && ((fmt == 1 && bits == 16) || (fmt == 3 && bits == 32))) {
                info.pcm16 = fmt == 1;
                void *dst = alloc(n * bps);
                rd(dst, n * bps, true);
                return info;
            }
            std::vector<unsigned char> raw(n * frame);
            rd(raw.data(), raw.size(), true);
            float *dst = static_cast<float *>(alloc(n * sizeof(float)));
            for (size_t i = 0; i < n; ++i) {
                double acc = 0.0;
                for (unsigned c = 0; c < channels; ++c) acc += detail::wav_sample(&raw[i * frame + c * bps], fmt, bits, filename);
                dst[i] = static_cast<float>(acc / channels);
            }
            return info;
        } else {
            pos += static_cast<off_t>(sz) + (sz & 1);
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

/// Writes mono 16-bit PCM WAV (what a decoder or a CD rip delivers; the bench's index()/search() inputs).
inline void write_wav_pcm16(const std::string &filename, const int16_t *samples, size_t n, int sample_rate) {
    std::ofstream os(filename, std::ios::binary);
    if (!os) throw std::runtime_error("cannot create '" + filename + "'");
    const uint32_t data_bytes = static_cast<uint32_t>(n * 2), riff = 36 + data_bytes, fmt_sz = 16, rate = sample_rate,
                   byte_rate = rate * 2;
    const uint16_t fmt = 1, ch = 1, align = 2, bits = 16;
    os.write("RIFF", 4); os.write(reinterpret_cast<const char *>(&riff), 4); os.write("WAVEfmt ", 8);
    os.write(reinterpret_cast<const char *>(&fmt_sz), 4); os.write(reinterpret_cast<const char *>(&fmt), 2);
    os.write(reinterpret_cast<const char *>(&ch), 2); os.write(reinterpret_cast<const char *>(&rate), 4);
    os.write(reinterpret_cast<const char *>(&byte_rate), 4); os.write(reinterpret_cast<const char *>(&align), 2);
    os.write(reinterpret_cast<const char *>(&bits), 2); os.write("data", 4);
    os.write(reinterpret_cast<const char *>(&data_bytes), 4);
    os.write(reinterpret_cast<const char *>(samples), data_bytes);
}

}  // namespace hpfw::io
