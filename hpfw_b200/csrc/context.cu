// hpfw_b200/csrc/context.cu — context lifetime, error reporting.
#include "common.cuh"

#include <cstdlib>
#include <cstring>

namespace hpfw_b200 {
static thread_local char g_err[1024] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace hpfw_b200

extern "C" {

const char *hpfw_last_error(void) { return hpfw_b200::g_err; }
const char *hpfw_version(void) { return "hpfw_b200 0.1 (sm_100a)"; }

int hpfw_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int hpfw_ctx_create(int device, hpfw_ctx **out) {
    if (!out) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_ctx_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        HPFW_FAIL(HPFW_ERR_CUDA, "hpfw_ctx_create: no CUDA device (%s); this library has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_ctx_create: device %d out of range [0,%d)", device, n);
    hpfw_b200::DeviceGuard g(device);
    cudaDeviceProp p;
    HPFW_CUDA_TRY(cudaGetDeviceProperties(&p, device));
    if (p.major < 10)
        HPFW_FAIL(HPFW_ERR_CUDA, "hpfw_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only", device,
                  p.major, p.minor);
    hpfw_ctx *c = new hpfw_ctx();
    c->device = device;
    c->sm_count = p.multiProcessorCount;
    c->max_smem_optin = int(p.sharedMemPerBlockOptin);
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->pin_in_free, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        delete c;
        HPFW_FAIL(HPFW_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e));
    }
    if (const char *env = getenv("HPFW_MATCH_IMPL")) {
        const int v = atoi(env);
        if (v >= 0 && v <= 3) c->match_impl = v;
    }
    if (const char *env = getenv("HPFW_MATCH_TC_F4")) c->match_tc_f4 = atoi(env) != 0;
    if (const char *env = getenv("HPFW_CQT_WINDOW")) {
        if (!strcmp(env, "symmetric") || !strcmp(env, "1")) c->cqt_window = 1;
        else if (!strcmp(env, "periodic") || !strcmp(env, "0")) c->cqt_window = 0;
        else {
            delete c;
            HPFW_FAIL(HPFW_ERR_ARG, "HPFW_CQT_WINDOW must be 'periodic' or 'symmetric' (got '%s')", env);
        }
    }
    *out = c;
    return HPFW_OK;
}

void hpfw_ctx_destroy(hpfw_ctx *c) {
    if (!c) return;
    hpfw_b200::DeviceGuard g(c->device);
    cudaStreamSynchronize(c->stream);
    c->best.release();
    c->qmeta.release();
    c->qwords.release();
    c->keys.release();
    c->qexp.release();
    c->pin_in.release();
    c->pin_out.release();
    c->filters_perm.release();
    c->spectro.release();
    c->hp.release();
    c->yproj.release();
    c->colmeta.release();
    c->filters_tc.release();
    c->filters_tc16.release();
    c->delta_tc.release();
    c->audio.release();
    c->audio_f.release();
    c->cov_accum.release();
    c->cov_scratch.release();
    c->cov_accum_side.release();
    c->cov_scratch_side.release();
    for (auto &b : c->eig_scratch) b.release();
    hpfw_b200::cqt_cache_destroy(c->cqt);
    for (auto &r : c->timing_pending) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    for (auto e : c->timing_pool) cudaEventDestroy(e);
    for (int l = 0; l < HPFW_CTX_LANES; ++l) {
        if (c->lane_stream[l]) cudaStreamDestroy(c->lane_stream[l]);
        if (c->lane_join[l]) cudaEventDestroy(c->lane_join[l]);
    }
    if (c->lane_fork) cudaEventDestroy(c->lane_fork);
    if (c->pin_in_free) cudaEventDestroy(c->pin_in_free);
    if (c->order_ev) cudaEventDestroy(c->order_ev);
    cudaStreamDestroy(c->stream);
    delete c;
}

int hpfw_ctx_device(const hpfw_ctx *c) { return c ? c->device : -1; }
uint64_t hpfw_ctx_launch_count(const hpfw_ctx *c) { return c ? c->launches : 0; }

int hpfw_ctx_timing_enable(hpfw_ctx *c, int on) {
    if (!c) HPFW_FAIL(HPFW_ERR_ARG, "ctx is NULL");
    c->timing = on != 0;
    return HPFW_OK;
}

int hpfw_ctx_timing_read(hpfw_ctx *c, int kernel, double *total_ms, uint64_t *launches, int reset) {
    if (!c || kernel < 0 || kernel >= HPFW_K_COUNT) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_ctx_timing_read: bad argument");
    hpfw_b200::DeviceGuard g(c->device);
    for (auto &r : c->timing_pending) {
        HPFW_CUDA_TRY(cudaEventSynchronize(r.b));
        float ms = 0.f;
        HPFW_CUDA_TRY(cudaEventElapsedTime(&ms, r.a, r.b));
        c->timing_ms[r.kernel] += double(ms);
        c->timing_n[r.kernel] += 1;
        c->timing_pool.push_back(r.a);
        c->timing_pool.push_back(r.b);
    }
    c->timing_pending.clear();
    if (total_ms) *total_ms = c->timing_ms[kernel];
    if (launches) *launches = c->timing_n[kernel];
    if (reset) {
        c->timing_ms[kernel] = 0.0;
        c->timing_n[kernel] = 0;
    }
    return HPFW_OK;
}

int hpfw_ctx_synchronize(hpfw_ctx *c) {
    if (!c) HPFW_FAIL(HPFW_ERR_ARG, "ctx is NULL");
    hpfw_b200::DeviceGuard g(c->device);
    HPFW_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return HPFW_OK;
}

}  // extern "C"
