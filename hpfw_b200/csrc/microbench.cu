// hpfw_b200/csrc/microbench.cu — register-only pipe microbenchmarks that pin the matcher's roofline denominator
// ("popc roof = SMs x 16 POPC.32/clk / 2 x f_SM — verify rate & clock by microbenchmark").
#include "common.cuh"

namespace hpfw_b200 {

constexpr int MB_CHAINS = 16;
constexpr int MB_UNROLL = 16;

// MODE 0: POPC only. MODE 1: 3-input LOP3 only. MODE 2: the matcher's inner mix (per 64-bit word-op: 2 LOP3, 2 POPC, 1 IADD3).
template <int MODE>
__global__ void __launch_bounds__(256) pipe_kernel(uint32_t *sink, int iters, unsigned long long *cycles) {
    uint32_t x[MB_CHAINS];
#pragma unroll
    for (int i = 0; i < MB_CHAINS; ++i) x[i] = sink[(threadIdx.x + i * 7) & 255] | 1u;
    uint32_t qx = sink[threadIdx.x & 31], qy = sink[(threadIdx.x + 3) & 31];
    __syncthreads();
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < MB_UNROLL; ++u) {
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < MB_CHAINS; ++i) x[i] = __popc(x[i]) | 0x10000u;
            } else if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < MB_CHAINS; ++i) x[i] = (x[i] ^ qx) & (x[(i + 1) % MB_CHAINS] | qy);
            } else {
                // 8 "reference words" (x[0..15] as lo/hi pairs are reused as the window), 8 accumulators in place
#pragma unroll
                for (int i = 0; i < MB_CHAINS / 2; ++i)
                    x[i] += __popc(qx ^ x[MB_CHAINS / 2 + i]) + __popc(qy ^ x[MB_CHAINS / 2 + ((i + 1) & 7)]);
                qx = qx * 1664525u + 1013904223u;
                qy ^= qx;
            }
        }
    }
    const unsigned long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < MB_CHAINS; ++i) s ^= x[i];
    if (s == 0x12345678u) sink[0] = s;  // keep the chains alive
    if (threadIdx.x == 0) atomicMax(cycles, t1 - t0);
}

}  // namespace hpfw_b200

using namespace hpfw_b200;

extern "C" int hpfw_microbench_pipes(hpfw_ctx *ctx, double *out, double *sm_clock_mhz_out) {
    if (!ctx || !out) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_microbench_pipes: NULL argument");
    DeviceGuard g(ctx->device);
    uint32_t *sink = nullptr;
    unsigned long long *cyc = nullptr;
    HPFW_CUDA_TRY(cudaMalloc(&sink, 256 * sizeof(uint32_t)));
    HPFW_CUDA_TRY(cudaMalloc(&cyc, sizeof(unsigned long long)));
    HPFW_CUDA_TRY(cudaMemset(sink, 0x5A, 256 * sizeof(uint32_t)));
    cudaEvent_t e0, e1;
    HPFW_CUDA_TRY(cudaEventCreate(&e0));
    HPFW_CUDA_TRY(cudaEventCreate(&e1));
    const int ctas_per_sm = 4, iters = 4000;
    const dim3 grid(ctx->sm_count * ctas_per_sm);
    double clk_mhz = 0.0;
    for (int mode = 0; mode < 3; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {  // rep 0 warms up
            HPFW_CUDA_TRY(cudaMemsetAsync(cyc, 0, sizeof(unsigned long long), ctx->stream));
            HPFW_CUDA_TRY(cudaEventRecord(e0, ctx->stream));
            if (mode == 0) pipe_kernel<0><<<grid, 256, 0, ctx->stream>>>(sink, iters, cyc);
            if (mode == 1) pipe_kernel<1><<<grid, 256, 0, ctx->stream>>>(sink, iters, cyc);
            if (mode == 2) pipe_kernel<2><<<grid, 256, 0, ctx->stream>>>(sink, iters, cyc);
            HPFW_CUDA_TRY(cudaGetLastError());
            ctx->launches++;  // (not a product kernel class: counted, never event-timed by KernelScope)
            HPFW_CUDA_TRY(cudaEventRecord(e1, ctx->stream));
            HPFW_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        }
        unsigned long long c = 0;
        HPFW_CUDA_TRY(cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost));
        float ms = 0.f;
        HPFW_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        // lane-ops per clock per SM, counting the op of interest
        const double threads_per_sm = 256.0 * ctas_per_sm;
        double ops_per_thread;
        if (mode == 0) ops_per_thread = double(iters) * MB_UNROLL * MB_CHAINS;            // POPC
        else if (mode == 1) ops_per_thread = double(iters) * MB_UNROLL * MB_CHAINS * 2;   // 2 LOP3 per chain step
        else ops_per_thread = double(iters) * MB_UNROLL * (MB_CHAINS / 2);                // 64-bit word-ops
        out[mode] = ops_per_thread * threads_per_sm / double(c);
        if (mode == 2) clk_mhz = double(c) / (double(ms) * 1e3);
    }
    if (sm_clock_mhz_out) *sm_clock_mhz_out = clk_mhz;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    cudaFree(cyc);
    return HPFW_OK;
}
