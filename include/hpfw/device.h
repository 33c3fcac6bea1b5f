// include/hpfw/device.h — C++ RAII layer over the C ABI (include/hpfw_b200.h). Header-only, like the reference library.
// Everything below the ABI is hand-written sm_100a CUDA in libhpfw_b200.so; there is no CPU fallback: constructing a
// Context on a machine without a B200-class GPU throws hpfw::Error.
#pragma once

#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>

#include "../hpfw_b200.h"

namespace hpfw {

/// Exception carrying the ABI status code (HPFW_ERR_*); the reference reports every failure as a std::exception too
/// (caught per file in ParallelCollector::preprocess and LiveSongIdentification::search).
class Error : public std::runtime_error {
public:
    Error(int code, const std::string &what) : std::runtime_error(what), code_(code) {}
    int code() const { return code_; }

private:
    int code_;
};

namespace device {

inline void check(int status) {
    if (status != HPFW_OK) throw Error(status, std::string("hpfw_b200: ") + hpfw_last_error());
}

/// One hpfw_ctx (CUDA device + stream + scratch). Not copyable; share through `shared(device)`.
class Context {
public:
    explicit Context(int device = 0) { check(hpfw_ctx_create(device, &ctx_)); }
    ~Context() { hpfw_ctx_destroy(ctx_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    hpfw_ctx *get() const { return ctx_; }
    std::mutex &mutex() { return mtx_; }   // the ABI context is not re-entrant: one call at a time per context

    /// Process-wide context of a device (created on first use).
    static std::shared_ptr<Context> shared(int device = 0) {
        static std::mutex m;
        static std::map<int, std::weak_ptr<Context>> table;
        std::scoped_lock l(m);
        auto sp = table[device].lock();
        if (!sp) {
            sp = std::make_shared<Context>(device);
            table[device] = sp;
        }
        return sp;
    }

    /// Whose filters the device context currently holds (a collector's id and the generation of its filter set): collectors
    /// that share a context re-upload only when the tag differs. Read and written under mutex().
    uint64_t filters_tag = 0;

private:
    hpfw_ctx *ctx_ = nullptr;
    std::mutex mtx_;
};

/// RAII handle of an extraction stream (include/hpfw_b200.h, xstream.cu): the pinned staging ring the decode threads fill and
/// the device-side store of resident spectrograms and hashprints.
class ExtractionStream {
public:
    ExtractionStream(Context &ctx, int slots, size_t slot_bytes) { check(hpfw_xs_create(ctx.get(), slots, slot_bytes, &xs_)); }
    ~ExtractionStream() { hpfw_xs_destroy(xs_); }
    ExtractionStream(const ExtractionStream &) = delete;
    ExtractionStream &operator=(const ExtractionStream &) = delete;
    hpfw_xs *get() const { return xs_; }
    /// bumped by the owner whenever it resets the stream: handles to earlier contents (DeviceHashprints) compare it
    uint64_t generation = 0;

private:
    hpfw_xs *xs_ = nullptr;
};

/// Pinned host buffer (destination of hpfw_xs_fetch_spectrogram in the cache-writer threads).
class PinnedBuffer {
public:
    PinnedBuffer() = default;
    ~PinnedBuffer() { hpfw_host_free(p_); }
    PinnedBuffer(const PinnedBuffer &) = delete;
    PinnedBuffer &operator=(const PinnedBuffer &) = delete;
    void *reserve(size_t bytes) {
        if (bytes > cap_) {
            hpfw_host_free(p_);
            p_ = nullptr;
            cap_ = 0;
            check(hpfw_host_alloc(bytes + bytes / 4, &p_));
            cap_ = bytes + bytes / 4;
        }
        return p_;
    }

private:
    void *p_ = nullptr;
    size_t cap_ = 0;
};

}  // namespace device
}  // namespace hpfw
