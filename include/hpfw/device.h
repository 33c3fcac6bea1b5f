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

private:
    hpfw_ctx *ctx_ = nullptr;
    std::mutex mtx_;
};

}  // namespace device
}  // namespace hpfw
