// hpfw_b200/csrc/project.cu — stages 2+3: context stacking, learned projection, temporal delta, threshold, bit packing.
//
// Replaces (paths relative to /root/reference/include/hpfw/):
//   HashprintHandle::calc_frames            core/hashprint_handle.h:79-93    (never materialised here)
//   `filters * frames`                      core/parallel_collector.h:57,127
//   HashprintHandle::calc_fingerprint       core/hashprint_handle.h:115-125
//   HashprintHandle::fingerprint_to_hashprint / bool_col_to_num   core/hashprint_handle.h:127-142
//
//   y[f,t]  = sum_{b<121} sum_{c<20} F[f, b*20+c] * S[b, t+c]
//   bit(f,t) = (y[f,t] - y[f,t+80]) >= 0,   hp[t] = sum_f bit(f,t) << (63-f)
//
// Fused "delta-first" formulation: by linearity y[f,t]-y[f,t+80] = sum F[f,.] * (S[., t+c] - S[., t+80+c]); the CTA builds the
// differenced tile D = S[:, t0..] - S[:, t0+80..] in shared memory once (band-major, so a thread's 8 frames are contiguous),
// streams the filters through a double-buffered cp.async ring, and thresholds/packs in the epilogue. Frames are a 20-tap
// sliding window over D, so each thread keeps a 27-float register window per band and reuses it for 4 filters x 20 taps:
// 640 FFMA per 27 LDS.128 — FMA-pipe bound. (A tcgen05 kind::tf32 variant is the planned replacement; see DESIGN.md.)
#include "common.cuh"

#include <cstring>

namespace hpfw_b200 {

constexpr int PJ_BINS = HPFW_BINS;        // 121
constexpr int PJ_CTX = HPFW_CONTEXT;      // 20
constexpr int PJ_LAG = HPFW_LAG;          // 80
constexpr int PJ_NF = HPFW_NFILTERS;      // 64
constexpr int PJ_THREADS = 256;
constexpr int PJ_TT = 128;                // frames (hashprint words) per CTA tile
constexpr int PJ_PITCH = 148;             // columns held per band row: TT + 19 rounded so the 28-float window stays inside
constexpr int PJ_BC = 4;                  // bands per filter chunk
constexpr int PJ_NCHUNK = (PJ_BINS + PJ_BC - 1) / PJ_BC;    // 31
constexpr int PJ_BINS_PAD = PJ_NCHUNK * PJ_BC;               // 124
constexpr int PJ_CHUNK_FLOATS = PJ_BC * PJ_NF * PJ_CTX;      // 5120 floats = 20 KB
constexpr size_t PJ_SMEM = sizeof(float) * (size_t(PJ_BINS) * PJ_PITCH + 2 * PJ_CHUNK_FLOATS);

struct ProjTile {
    int32_t track;
    int32_t t0;
};

struct ProjTrack {
    int64_t col_start;   // first spectrogram column of this track in the concatenated buffer
    int64_t out_start;   // first output element (hashprint word, or y column) of this track
    int32_t cols;
    int32_t n_out;       // words (cols-99) or frames (cols-19) to write
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const uint32_t s = uint32_t(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// filters_perm: [band (124)][j (4)][fg (16)][c (20)] for filter f = fg*4 + j; zero for band >= 121.
// MODE 0: hashprint words (delta-first, threshold, pack). MODE 1: y itself (no delta), column-major [64 x frames].
template <int MODE>
__global__ void __launch_bounds__(PJ_THREADS, 2)
project_kernel(const float *__restrict__ spectro, const float *__restrict__ filters_perm,
               const ProjTile *__restrict__ tiles, const ProjTrack *__restrict__ tracks,
               uint64_t *__restrict__ hp_out, float *__restrict__ y_out) {
    extern __shared__ __align__(16) float psm[];
    float *Dt = psm;                                  // [121][PJ_PITCH]
    float *Fs = psm + PJ_BINS * PJ_PITCH;             // 2 x [4][4][16][20]

    const int tid = threadIdx.x;
    const ProjTile tile = tiles[blockIdx.x];
    const ProjTrack trk = tracks[tile.track];
    const float *S = spectro + trk.col_start * PJ_BINS;
    const int t0 = tile.t0;

    // prefetch filter chunk 0
    {
        const float4 *src = reinterpret_cast<const float4 *>(filters_perm);
        float4 *dst = reinterpret_cast<float4 *>(Fs);
        for (int i = tid; i < PJ_CHUNK_FLOATS / 4; i += PJ_THREADS) cp_async16(dst + i, src + i);
        cp_async_commit();
    }
    // differenced (or plain) tile, transposed to band-major
    for (int e = tid; e < PJ_PITCH * PJ_BINS; e += PJ_THREADS) {
        const int col = e / PJ_BINS, b = e - col * PJ_BINS;
        const int t = t0 + col;
        float v = 0.f;
        if (t < trk.cols) {
            v = S[size_t(t) * PJ_BINS + b];
            if (MODE == 0) v -= (t + PJ_LAG < trk.cols) ? S[size_t(t + PJ_LAG) * PJ_BINS + b] : 0.f;
        }
        Dt[b * PJ_PITCH + col] = v;
    }

    const int fg = tid & 15, tg = tid >> 4;
    float acc[4][8];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;

    for (int ch = 0; ch < PJ_NCHUNK; ++ch) {
        // prefetch the next chunk into the other buffer, then wait for the current one
        if (ch + 1 < PJ_NCHUNK) {
            const float4 *src = reinterpret_cast<const float4 *>(filters_perm + size_t(ch + 1) * PJ_CHUNK_FLOATS);
            float4 *dst = reinterpret_cast<float4 *>(Fs + ((ch + 1) & 1) * PJ_CHUNK_FLOATS);
            for (int i = tid; i < PJ_CHUNK_FLOATS / 4; i += PJ_THREADS) cp_async16(dst + i, src + i);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();   // chunk ch (and, first time, Dt) visible to all
        const float *Fc = Fs + (ch & 1) * PJ_CHUNK_FLOATS;
        const int nb = min(PJ_BC, PJ_BINS - ch * PJ_BC);
        for (int bl = 0; bl < nb; ++bl) {
            const int b = ch * PJ_BC + bl;
            float d[28];
            const float4 *dp = reinterpret_cast<const float4 *>(Dt + b * PJ_PITCH + tg * 8);
#pragma unroll
            for (int v = 0; v < 7; ++v) {
                const float4 x = dp[v];
                d[4 * v] = x.x; d[4 * v + 1] = x.y; d[4 * v + 2] = x.z; d[4 * v + 3] = x.w;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float a[20];
                const float4 *ap = reinterpret_cast<const float4 *>(Fc + ((bl * 4 + j) * 16 + fg) * PJ_CTX);
#pragma unroll
                for (int v = 0; v < 5; ++v) {
                    const float4 x = ap[v];
                    a[4 * v] = x.x; a[4 * v + 1] = x.y; a[4 * v + 2] = x.z; a[4 * v + 3] = x.w;
                }
#pragma unroll
                for (int c = 0; c < PJ_CTX; ++c)
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[j][i] = fmaf(a[c], d[c + i], acc[j][i]);
            }
        }
        __syncthreads();   // everyone done with buffer (ch&1) before it is refilled at ch+2
    }

    if (MODE == 0) {
        // threshold (delta >= 0 -> 1, hashprint_handle.h:121) and pack: filter f -> bit 63-f (filter 0 = MSB)
        uint64_t mine = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t nib = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) nib |= (acc[j][i] >= 0.f ? 1u : 0u) << (3 - j);
            uint64_t w = uint64_t(nib) << (60 - 4 * fg);
#pragma unroll
            for (int s = 8; s > 0; s >>= 1) w |= __shfl_xor_sync(0xFFFFFFFFu, w, s);   // OR over the 16 filter groups
            if (fg == i) mine = w;
        }
        if (fg < 8) {
            const int t = t0 + tg * 8 + fg;
            if (t < trk.n_out) hp_out[trk.out_start + t] = mine;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int t = t0 + tg * 8 + i;
            if (t < trk.n_out) {
                float4 v = make_float4(acc[0][i], acc[1][i], acc[2][i], acc[3][i]);
                *reinterpret_cast<float4 *>(y_out + (size_t(trk.out_start) + t) * PJ_NF + fg * 4) = v;
            }
        }
    }
}

}  // namespace hpfw_b200

using namespace hpfw_b200;

static int run_project(hpfw_ctx *ctx, int mode, const float *d_spectro, const int64_t *col_offsets, int n,
                       uint64_t *d_hp, float *d_y, cudaStream_t stream) {
    if (!ctx->have_filters) HPFW_FAIL(HPFW_ERR_STATE, "filters not set: call hpfw_set_filters first");
    std::vector<ProjTrack> tracks(size_t(std::max(n, 1)));
    std::vector<ProjTile> tiles;
    int64_t out = 0;
    for (int i = 0; i < n; ++i) {
        const int64_t cols = col_offsets[i + 1] - col_offsets[i];
        if (cols < 0 || cols > (int64_t(1) << 30)) HPFW_FAIL(HPFW_ERR_ARG, "bad column count for spectrogram %d", i);
        const int64_t n_out = std::max<int64_t>(0, mode == 0 ? cols - (PJ_CTX - 1) - PJ_LAG : cols - (PJ_CTX - 1));
        tracks[i] = {col_offsets[i], out, int32_t(cols), int32_t(n_out)};
        for (int64_t t0 = 0; t0 < n_out; t0 += PJ_TT) tiles.push_back({i, int32_t(t0)});
        out += n_out;
    }
    if (tiles.empty()) return HPFW_OK;
    const size_t tb = sizeof(ProjTrack) * tracks.size(), lb = sizeof(ProjTile) * tiles.size();
    HPFW_CUDA_TRY(cudaEventSynchronize(ctx->pin_in_free));
    HPFW_TRY(ctx->pin_in.reserve(tb + lb));
    HPFW_TRY(ctx->colmeta.reserve(tb + lb));
    memcpy(ctx->pin_in.ptr, tracks.data(), tb);
    memcpy(ctx->pin_in.as<char>() + tb, tiles.data(), lb);
    HPFW_CUDA_TRY(cudaMemcpyAsync(ctx->colmeta.ptr, ctx->pin_in.ptr, tb + lb, cudaMemcpyHostToDevice, stream));
    HPFW_CUDA_TRY(cudaEventRecord(ctx->pin_in_free, stream));
    const ProjTrack *d_tracks = ctx->colmeta.as<ProjTrack>();
    const ProjTile *d_tiles = reinterpret_cast<const ProjTile *>(ctx->colmeta.as<char>() + tb);
    KernelScope ks(ctx, HPFW_K_PROJECT, stream);
    if (mode == 0) {
        HPFW_CUDA_TRY(cudaFuncSetAttribute(project_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(PJ_SMEM)));
        project_kernel<0><<<unsigned(tiles.size()), PJ_THREADS, PJ_SMEM, stream>>>(
            d_spectro, ctx->filters_perm.as<float>(), d_tiles, d_tracks, d_hp, nullptr);
    } else {
        HPFW_CUDA_TRY(cudaFuncSetAttribute(project_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(PJ_SMEM)));
        project_kernel<1><<<unsigned(tiles.size()), PJ_THREADS, PJ_SMEM, stream>>>(
            d_spectro, ctx->filters_perm.as<float>(), d_tiles, d_tracks, nullptr, d_y);
    }
    HPFW_CUDA_TRY(cudaGetLastError());
    return HPFW_OK;
}

extern "C" {

int hpfw_set_filters(hpfw_ctx *ctx, const float *f) {
    if (!ctx || !f) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_set_filters: NULL argument");
    DeviceGuard g(ctx->device);
    ctx->filters_host.assign(f, f + size_t(PJ_NF) * HPFW_FRAME_SIZE);
    // [band 124][j 4][fg 16][c 20], filter = fg*4 + j; reference element F(f, b*20+c) sits at f + 64*(b*20+c)
    std::vector<float> perm(size_t(PJ_BINS_PAD) * PJ_NF * PJ_CTX, 0.f);
    for (int b = 0; b < PJ_BINS; ++b)
        for (int j = 0; j < 4; ++j)
            for (int fgi = 0; fgi < 16; ++fgi)
                for (int c = 0; c < PJ_CTX; ++c)
                    perm[((size_t(b) * 4 + j) * 16 + fgi) * PJ_CTX + c] = f[(fgi * 4 + j) + size_t(PJ_NF) * (b * PJ_CTX + c)];
    HPFW_TRY(ctx->filters_perm.reserve(sizeof(float) * perm.size()));
    HPFW_CUDA_TRY(cudaMemcpy(ctx->filters_perm.ptr, perm.data(), sizeof(float) * perm.size(), cudaMemcpyHostToDevice));
    HPFW_TRY(project_tc_set_filters(ctx, f));
    ctx->have_filters = true;
    return HPFW_OK;
}

int hpfw_get_filters(hpfw_ctx *ctx, float *f) {
    if (!ctx || !f) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_get_filters: NULL argument");
    if (!ctx->have_filters) HPFW_FAIL(HPFW_ERR_STATE, "filters not set");
    memcpy(f, ctx->filters_host.data(), sizeof(float) * ctx->filters_host.size());
    return HPFW_OK;
}

int hpfw_hashprint_words_for_cols(int cols) { return cols - (PJ_CTX - 1) - PJ_LAG; }

int hpfw_hashprint_from_spectrogram_device(hpfw_ctx *ctx, const float *d_spectrograms, const int64_t *col_offsets, int n,
                                           uint64_t *d_hp_out, void *stream) {
    if (!ctx || !col_offsets || n < 0) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_hashprint_from_spectrogram_device: bad argument");
    if (n == 0) return HPFW_OK;
    if (!d_spectrograms || !d_hp_out) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_hashprint_from_spectrogram_device: NULL buffer");
    DeviceGuard g(ctx->device);
    if (!ctx->have_filters) HPFW_FAIL(HPFW_ERR_STATE, "filters not set: call hpfw_set_filters first");
    if (ctx->project_impl != 0)
        return project_tc_run(ctx, ctx->project_impl, d_spectrograms, col_offsets, n, d_hp_out, ctx->pick(stream));
    return run_project(ctx, 0, d_spectrograms, col_offsets, n, d_hp_out, nullptr, ctx->pick(stream));
}

int hpfw_set_projection_impl(hpfw_ctx *ctx, int impl) {
    if (!ctx || impl < 0 || impl > 5) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_set_projection_impl: impl must be 0 .. 5");
    ctx->project_impl = impl;
    return HPFW_OK;
}

int hpfw_hashprint_from_spectrogram(hpfw_ctx *ctx, const float *spectrogram, int cols, uint64_t *hp_out, int *n_out) {
    if (!ctx || !spectrogram || !hp_out || !n_out) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_hashprint_from_spectrogram: NULL argument");
    *n_out = 0;
    const int n = hpfw_hashprint_words_for_cols(cols);
    // the reference underflows size_t here (hashprint_handle.h:84,119) and dies with bad_alloc; we report it
    if (n <= 0) HPFW_FAIL(HPFW_ERR_SHORT, "spectrogram has %d columns; at least 100 are needed for one hashprint word", cols);
    DeviceGuard g(ctx->device);
    ctx->order_on(ctx->stream);
    const size_t sb = sizeof(float) * size_t(cols) * PJ_BINS;
    HPFW_TRY(ctx->spectro.reserve(sb));
    HPFW_TRY(ctx->hp.reserve(sizeof(uint64_t) * size_t(n)));
    HPFW_CUDA_TRY(cudaMemcpyAsync(ctx->spectro.ptr, spectrogram, sb, cudaMemcpyHostToDevice, ctx->stream));
    const int64_t co[2] = {0, cols};
    HPFW_TRY(hpfw_hashprint_from_spectrogram_device(ctx, ctx->spectro.as<float>(), co, 1, ctx->hp.as<uint64_t>(), ctx->stream));
    HPFW_CUDA_TRY(cudaMemcpyAsync(hp_out, ctx->hp.ptr, sizeof(uint64_t) * size_t(n), cudaMemcpyDeviceToHost, ctx->stream));
    HPFW_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    *n_out = n;
    return HPFW_OK;
}

int hpfw_project(hpfw_ctx *ctx, const float *spectrogram, int cols, float *y_out) {
    if (!ctx || !spectrogram || !y_out) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_project: NULL argument");
    const int nfr = cols - (PJ_CTX - 1);
    if (nfr <= 0) HPFW_FAIL(HPFW_ERR_SHORT, "spectrogram has %d columns; at least 20 are needed for one frame", cols);
    DeviceGuard g(ctx->device);
    ctx->order_on(ctx->stream);
    const size_t sb = sizeof(float) * size_t(cols) * PJ_BINS, yb = sizeof(float) * size_t(nfr) * PJ_NF;
    HPFW_TRY(ctx->spectro.reserve(sb));
    HPFW_TRY(ctx->yproj.reserve(yb));
    HPFW_CUDA_TRY(cudaMemcpyAsync(ctx->spectro.ptr, spectrogram, sb, cudaMemcpyHostToDevice, ctx->stream));
    const int64_t co[2] = {0, cols};
    HPFW_TRY(run_project(ctx, 1, ctx->spectro.as<float>(), co, 1, nullptr, ctx->yproj.as<float>(), ctx->stream));
    HPFW_CUDA_TRY(cudaMemcpyAsync(y_out, ctx->yproj.ptr, yb, cudaMemcpyDeviceToHost, ctx->stream));
    HPFW_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return HPFW_OK;
}

}  // extern "C"
