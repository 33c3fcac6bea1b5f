// hpfw_b200/csrc/pcm.cu — 16-bit PCM entry points of the query / index path.
//
// The reference decodes every file to float on the CPU (essentia MonoLoader, /root/reference/include/hpfw/spectrum/cqt.h:45-52)
// and the float entry points of this library take that buffer. A 3-minute track is 31.75 MB of floats, and from host memory
// the extraction is bound by the PCIe copy, not by the kernels; the same audio as 16-bit PCM — what a WAV file or a decoder
// delivers before the conversion — is half the bytes. These entry points copy the PCM samples and do MonoLoader's
// `sample / 32768` on the device (exact in float, so the hashprints are those of the float path bit for bit).
#include "common.cuh"

#include <algorithm>
#include <vector>

namespace hpfw_b200 {

// 8 samples per thread: one 16-byte load, two 16-byte stores
__global__ void __launch_bounds__(256)
pcm16_to_float_kernel(const int16_t *__restrict__ in, float *__restrict__ out, int64_t n) {
    const int64_t i8 = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
    if (i8 >= n) return;
    constexpr float S = 1.0f / 32768.0f;
    if (i8 + 8 <= n && ((reinterpret_cast<uintptr_t>(in + i8) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out + i8) & 15) == 0)) {
        const uint4 v = *reinterpret_cast<const uint4 *>(in + i8);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        float f[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            f[2 * j] = float(int16_t(w[j] & 0xFFFFu)) * S;
            f[2 * j + 1] = float(int16_t(w[j] >> 16)) * S;
        }
        reinterpret_cast<float4 *>(out + i8)[0] = make_float4(f[0], f[1], f[2], f[3]);
        reinterpret_cast<float4 *>(out + i8)[1] = make_float4(f[4], f[5], f[6], f[7]);
    } else {
        for (int64_t i = i8; i < std::min<int64_t>(n, i8 + 8); ++i) out[i] = float(in[i]) * S;
    }
}

int pcm16_convert(hpfw_ctx *ctx, const int16_t *d_pcm, float *d_out, int64_t n, cudaStream_t s) {
    if (n <= 0) return HPFW_OK;
    KernelScope ks(ctx, HPFW_K_OTHER, s);
    const int64_t threads = (n + 7) / 8;
    pcm16_to_float_kernel<<<unsigned((threads + 255) / 256), 256, 0, s>>>(d_pcm, d_out, n);
    HPFW_CUDA_TRY(cudaGetLastError());
    return HPFW_OK;
}

}  // namespace hpfw_b200

using namespace hpfw_b200;

extern "C" {

int hpfw_pcm16_to_float_device(hpfw_ctx *ctx, const int16_t *d_pcm, int64_t n_samples, float *d_audio_out, void *stream) {
    if (!ctx || n_samples < 0 || (n_samples > 0 && (!d_pcm || !d_audio_out)))
        HPFW_FAIL(HPFW_ERR_ARG, "hpfw_pcm16_to_float_device: bad argument");
    DeviceGuard g(ctx->device);
    return pcm16_convert(ctx, d_pcm, d_audio_out, n_samples, ctx->pick(stream));
}

int hpfw_cqt_spectrogram_pcm16(hpfw_ctx *ctx, const int16_t *pcm, int64_t n_samples, float *spectrogram_out, int *cols_out) {
    if (!ctx || !pcm || !spectrogram_out || !cols_out) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_cqt_spectrogram_pcm16: NULL argument");
    *cols_out = 0;
    DeviceGuard g(ctx->device);
    ctx->order_on(ctx->stream);
    const int cols = hpfw_cqt_cols(n_samples);
    if (cols <= 0) HPFW_FAIL(HPFW_ERR_SHORT, "audio of %lld samples is too short for the CQT design", (long long)n_samples);
    HPFW_TRY(ctx->audio.reserve(sizeof(int16_t) * size_t(n_samples + 8)));
    HPFW_TRY(ctx->audio_f.reserve(sizeof(float) * size_t(n_samples + 8)));
    HPFW_TRY(ctx->spectro.reserve(sizeof(float) * size_t(cols) * HPFW_BINS));
    HPFW_CUDA_TRY(cudaMemcpyAsync(ctx->audio.ptr, pcm, sizeof(int16_t) * size_t(n_samples), cudaMemcpyHostToDevice,
                                  ctx->stream));
    HPFW_TRY(pcm16_convert(ctx, ctx->audio.as<int16_t>(), ctx->audio_f.as<float>(), n_samples, ctx->stream));
    HPFW_TRY(hpfw_cqt_spectrogram_device(ctx, ctx->audio_f.as<float>(), n_samples, ctx->spectro.as<float>(), ctx->stream));
    HPFW_CUDA_TRY(cudaMemcpyAsync(spectrogram_out, ctx->spectro.ptr, sizeof(float) * size_t(cols) * HPFW_BINS,
                                  cudaMemcpyDeviceToHost, ctx->stream));
    HPFW_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    *cols_out = cols;
    return HPFW_OK;
}

int hpfw_calc_hashprint_pcm16_batch_device(hpfw_ctx *ctx, const int16_t *d_pcm, const int64_t *sample_offsets, int n,
                                           uint64_t *d_hp_out, void *stream) {
    if (!ctx || !sample_offsets || n < 0) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_calc_hashprint_pcm16_batch_device: bad argument");
    if (n == 0) return HPFW_OK;
    if (!d_pcm || !d_hp_out) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_calc_hashprint_pcm16_batch_device: NULL buffer");
    DeviceGuard g(ctx->device);
    cudaStream_t s = ctx->pick(stream);
    // chunks of tracks whose float copy stays below 1 GiB
    const int64_t chunk_samples = int64_t(1) << 28;
    int i = 0;
    int64_t hp_off = 0;
    while (i < n) {
        int j = i + 1;
        while (j < n && sample_offsets[j + 1] - sample_offsets[i] <= chunk_samples) ++j;
        const int64_t base = sample_offsets[i], total = sample_offsets[j] - base;
        if (total < 0) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_calc_hashprint_pcm16_batch_device: offsets must be monotone");
        HPFW_TRY(ctx->audio_f.reserve(sizeof(float) * size_t(total + 8)));
        HPFW_TRY(pcm16_convert(ctx, d_pcm + base, ctx->audio_f.as<float>(), total, s));
        std::vector<int64_t> rel(size_t(j - i) + 1);
        for (int t = i; t <= j; ++t) rel[t - i] = sample_offsets[t] - base;
        HPFW_TRY(hpfw_calc_hashprint_audio_batch_device(ctx, ctx->audio_f.as<float>(), rel.data(), j - i, d_hp_out + hp_off, s));
        for (int t = i; t < j; ++t) hp_off += std::max(0, hpfw_hashprint_words_for_samples(rel[t - i + 1] - rel[t - i]));
        i = j;
    }
    return HPFW_OK;
}

int hpfw_calc_hashprint_pcm16_device(hpfw_ctx *ctx, const int16_t *d_pcm, int64_t n_samples, uint64_t *d_hp_out,
                                     void *stream) {
    if (!ctx || !d_pcm || !d_hp_out) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_calc_hashprint_pcm16_device: NULL argument");
    DeviceGuard g(ctx->device);
    cudaStream_t s = ctx->pick(stream);
    if (hpfw_hashprint_words_for_samples(n_samples) <= 0)
        HPFW_FAIL(HPFW_ERR_SHORT, "audio of %lld samples is too short for one hashprint word", (long long)n_samples);
    HPFW_TRY(ctx->audio_f.reserve(sizeof(float) * size_t(n_samples + 8)));
    HPFW_TRY(pcm16_convert(ctx, d_pcm, ctx->audio_f.as<float>(), n_samples, s));
    return hpfw_calc_hashprint_audio_device(ctx, ctx->audio_f.as<float>(), n_samples, d_hp_out, s);
}

int hpfw_calc_hashprint_pcm16(hpfw_ctx *ctx, const int16_t *pcm, int64_t n_samples, uint64_t *hp_out, int *n_out) {
    if (!ctx || !pcm || !hp_out || !n_out) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_calc_hashprint_pcm16: NULL argument");
    *n_out = 0;
    DeviceGuard g(ctx->device);
    ctx->order_on(ctx->stream);
    const int n = hpfw_hashprint_words_for_samples(n_samples);
    if (n <= 0)
        HPFW_FAIL(HPFW_ERR_SHORT, "audio of %lld samples is too short for one hashprint word", (long long)n_samples);
    HPFW_TRY(ctx->audio.reserve(sizeof(int16_t) * size_t(n_samples + 8)));
    HPFW_TRY(ctx->hp.reserve(sizeof(uint64_t) * size_t(n)));
    HPFW_CUDA_TRY(cudaMemcpyAsync(ctx->audio.ptr, pcm, sizeof(int16_t) * size_t(n_samples), cudaMemcpyHostToDevice,
                                  ctx->stream));
    HPFW_TRY(hpfw_calc_hashprint_pcm16_device(ctx, ctx->audio.as<int16_t>(), n_samples, ctx->hp.as<uint64_t>(), ctx->stream));
    HPFW_CUDA_TRY(cudaMemcpyAsync(hp_out, ctx->hp.ptr, sizeof(uint64_t) * size_t(n), cudaMemcpyDeviceToHost, ctx->stream));
    HPFW_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    *n_out = n;
    return HPFW_OK;
}

}  // extern "C"
