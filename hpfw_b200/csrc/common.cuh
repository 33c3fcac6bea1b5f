// hpfw_b200/csrc/common.cuh — shared host-side plumbing for the C ABI (include/hpfw_b200.h).
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/hpfw_b200.h"

namespace hpfw_b200 {

void set_error(const char *fmt, ...);

#define HPFW_CUDA_TRY(expr)                                                                          \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            hpfw_b200::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,                  \
                                 cudaGetErrorString(_e));                                            \
            return HPFW_ERR_CUDA;                                                                    \
        }                                                                                            \
    } while (0)

#define HPFW_TRY(expr)                   \
    do {                                 \
        int _s = (expr);                 \
        if (_s != HPFW_OK) return _s;    \
    } while (0)

#define HPFW_FAIL(code, ...)               \
    do {                                   \
        hpfw_b200::set_error(__VA_ARGS__); \
        return (code);                     \
    } while (0)

// Grow-only device scratch buffer.
struct DeviceBuffer {
    void *ptr = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return HPFW_OK;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        HPFW_CUDA_TRY(cudaMalloc(&ptr, want));
        cap = want;
        return HPFW_OK;
    }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
    }
    template <class T> T *as() const { return static_cast<T *>(ptr); }
};

// Grow-only pinned host staging buffer.
struct PinnedBuffer {
    void *ptr = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return HPFW_OK;
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        HPFW_CUDA_TRY(cudaMallocHost(&ptr, want));
        cap = want;
        return HPFW_OK;
    }
    void release() {
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr;
        cap = 0;
    }
    template <class T> T *as() const { return static_cast<T *>(ptr); }
};

struct CqtPlanCache;  // cqt.cu
}  // namespace hpfw_b200
#define HPFW_CTX_LANES 8
namespace hpfw_b200 {

}  // namespace hpfw_b200

struct hpfw_ctx {
    int device = 0;
    int sm_count = 0;
    int max_smem_optin = 0;
    cudaStream_t stream = nullptr;  // the context's own stream
    cudaEvent_t pin_in_free = nullptr;  // recorded after the last H2D copy out of pin_in
    uint64_t launches = 0;

    // matcher scratch
    hpfw_b200::DeviceBuffer best;      // [queries_in_chunk][tracks] u64
    hpfw_b200::DeviceBuffer qmeta;     // query block tables
    hpfw_b200::DeviceBuffer qwords;    // staged query words (host entry points)
    hpfw_b200::DeviceBuffer keys;      // staged result keys (host entry points)
    hpfw_b200::PinnedBuffer pin_in, pin_out;
    hpfw_b200::DeviceBuffer qexp;      // match_tc.cu: queries expanded to signed bytes
    int match_impl = 2;                // 2 = tensor cores for (nearly) full groups of 128 queries, integer pipes for the rest
                                       // (default); 1 = tensor cores (int8) always; 3 = tensor cores (fp4) always;
                                       // 0 = integer pipes always
    int match_tc_f4 = 1;               // operand encoding the default routing (impl 2) uses: 1 = fp4 (default), 0 = int8
    int tc_selftest[2] = {0, 0};       // known-answer self-test of the tensor-core matcher per encoding [int8, fp4]:
                                       // 0 = not run yet, 1 = passed, -1 = failed (route disabled), 2 = running

    // projection state
    hpfw_b200::DeviceBuffer filters_perm;  // filters permuted to [context][band(padded 128)][filter] etc. (project.cu)
    std::vector<float> filters_host;       // column-major 64 x 2420 as given
    bool have_filters = false;
    hpfw_b200::DeviceBuffer spectro, hp, yproj, colmeta;
    hpfw_b200::DeviceBuffer filters_tc16;           // project_tc.cu impl 3: fp16 filters [tap][filter][band]
    hpfw_b200::DeviceBuffer filters_tc, delta_tc;   // project_tc.cu: tf32 filters [tap][filter][band], differenced spectrogram
    int project_impl = 4;                           // 4 = tcgen05, fp16 operands, 2 tiles per CTA share the filter stream (default);
                                                    // 5 = the same with 4 tiles; 3 = one tile per CTA; 1 = tcgen05, tf32 operands;
                                                    // 2 = tcgen05, tf32, A window reloaded per tap; 0 = CUDA-core kernel

    // filter learning (learn.cu): device-resident covariance accumulator (2420 x 2420) and scratch
    hpfw_b200::DeviceBuffer cov_accum, cov_scratch;
    hpfw_b200::DeviceBuffer cov_accum_side, cov_scratch_side;   // second covariance slot of the extraction stream
    bool cov_side_dirty = false;
    hpfw_b200::DeviceBuffer eig_scratch[6];   // hpfw_calc_filters: iteration blocks, Gram / rotation matrices, residuals
    uint64_t cov_tracks = 0;

    // cqt state
    int cqt_window = 0;                // band window: 0 = periodic centred Hann (default), 1 = symmetric Hann (cqt.cu)
    hpfw_b200::CqtPlanCache *cqt = nullptr;
    cudaStream_t lane_stream[HPFW_CTX_LANES] = {};   // batched extraction: concurrent tracks
    cudaEvent_t lane_join[HPFW_CTX_LANES] = {};
    cudaEvent_t lane_fork = nullptr;
    hpfw_b200::DeviceBuffer audio;
    hpfw_b200::DeviceBuffer audio_f;   // pcm.cu: float copy of 16-bit PCM input

    // optional per-kernel device timing (CUDA events on the launching stream); see hpfw_ctx_timing_*
    bool timing = false;
    struct TimingRec { int kernel; cudaEvent_t a, b; };
    std::vector<TimingRec> timing_pending;
    std::vector<cudaEvent_t> timing_pool;
    double timing_ms[HPFW_K_COUNT] = {};
    uint64_t timing_n[HPFW_K_COUNT] = {};

    // Cross-stream ordering. All entry points of a context share its scratch buffers (best/qmeta/qexp, spectro, delta_tc,
    // cov_accum, the CQT lanes ...), so work enqueued on one stream must not overtake work an earlier call enqueued on
    // another. Every entry point announces the stream it is about to use: when that differs from the stream of the previous
    // call, the new stream first waits (on the device, no host synchronisation) for everything enqueued so far on the old one.
    // A caller that keeps to one stream per context pays nothing.
    cudaStream_t last_stream = nullptr;
    bool last_stream_valid = false;
    cudaEvent_t order_ev = nullptr;
    void order_on(cudaStream_t s) {
        if (last_stream_valid && last_stream != s) {
            if (!order_ev) cudaEventCreateWithFlags(&order_ev, cudaEventDisableTiming);
            // a caller may have destroyed the previous stream (its work is then complete or abandoned): ignore that error
            if (order_ev && cudaEventRecord(order_ev, last_stream) == cudaSuccess) cudaStreamWaitEvent(s, order_ev, 0);
            else cudaGetLastError();
        }
        last_stream = s;
        last_stream_valid = true;
    }
    cudaStream_t pick(void *s) {
        cudaStream_t st = s ? static_cast<cudaStream_t>(s) : stream;
        order_on(st);
        return st;
    }
};

namespace hpfw_b200 {
// RAII device guard: every ABI call runs on its context's device.
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
void cqt_cache_destroy(CqtPlanCache *);
// internal (no stream-ordering hook; xstream.cu forks and joins its own streams around these)
int ctx_lanes_init(hpfw_ctx *ctx);                 // creates ctx->lane_stream[] / lane_join[] / lane_fork on first use
int cqt_run_lane(hpfw_ctx *ctx, const float *d_audio, int64_t n_samples, float *d_out, cudaStream_t stream, int lane);
int pcm16_convert(hpfw_ctx *ctx, const int16_t *d_pcm, float *d_out, int64_t n, cudaStream_t s);
int cov_add_device(hpfw_ctx *ctx, const float *d_spec, int cols, cudaStream_t s, int slot = 0);
int cov_fold_side(hpfw_ctx *ctx, cudaStream_t s);      // main accumulator += side accumulator (slot 1), side = 0
int project_tc_set_filters(hpfw_ctx *ctx, const float *filters_colmajor);
int project_tc_run(hpfw_ctx *ctx, int impl, const float *d_spectro, const int64_t *col_offsets, int n, uint64_t *d_hp,
                   cudaStream_t stream);

// Scope guard around ONE kernel launch: counts it and, when timing is enabled, brackets it with two events.
struct KernelScope {
    hpfw_ctx *ctx;
    cudaStream_t stream;
    int kernel;
    cudaEvent_t a = nullptr, b = nullptr;
    static cudaEvent_t get(hpfw_ctx *c) {
        cudaEvent_t e = nullptr;
        if (!c->timing_pool.empty()) {
            e = c->timing_pool.back();
            c->timing_pool.pop_back();
        } else {
            cudaEventCreate(&e);
        }
        return e;
    }
    KernelScope(hpfw_ctx *c, int k, cudaStream_t s) : ctx(c), stream(s), kernel(k) {
        ctx->launches++;
        if (ctx->timing) {
            a = get(ctx);
            b = get(ctx);
            cudaEventRecord(a, stream);
        }
    }
    ~KernelScope() {
        if (a) {
            cudaEventRecord(b, stream);
            ctx->timing_pending.push_back({kernel, a, b});
        }
    }
};
}  // namespace hpfw_b200
