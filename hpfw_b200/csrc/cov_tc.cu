// hpfw_b200/csrc/cov_tc.cu — the correlation GEMMs of calc_cov on the 5th-generation tensor cores.
//
// HashprintHandle::calc_cov (/root/reference/include/hpfw/core/hashprint_handle.h:96-102) over context frames reduces to
// twenty time-lagged correlations of the centred spectrogram (learn.cu header):
//     T0_d[b, b'] = sum_{u < nf} Sc[u][b] * Sc[u + d][b'],   d = 0 .. 19,   b, b' < 121 (padded to 128).
// Per lag that is a 128 x 128 x nf GEMM whose contraction runs over TIME, i.e. over the rows of the [cols][128] matrix:
// both operands are MN-major (the 128 bands of a row are contiguous) and the B operand of lag d is the A operand's block
// shifted down by d rows. One [rows][128] block per pipeline stage is loaded by TMA (4 column blocks of 32 bands) and every
// lag addresses it through the shared-memory descriptor's start address (+ d * 128 bytes), exactly as the projection kernel
// does for its context taps; no lagged copy is ever made.
//
// Precision: tf32 inputs would give the covariance 10 mantissa bits, so each value is split x = hi + lo (both tf32) by the
// centring pass and a product is three MMAs, hi*hi + hi*lo + lo*hi (the dropped lo*lo term is 2^-22 relative), accumulated in
// fp32 in TMEM: fp32-grade results at a third of the tf32 rate, still ~15x the CUDA-core kernel it replaces
// (t0_gemm_kernel, kept as implementation 0).
//
// Work split: a CTA owns one group of 4 lags (4 accumulators x 128 TMEM columns = all 512) and a contiguous range of
// 8-row K steps; it streams its range in chunks of CV_CHUNK rows through a 2-stage TMA/mbarrier ring and writes its four
// 128 x 128 partial sums once. Grid = (time splits, 5 lag groups). The <= 7 rows beyond the last whole K step are added by
// cov_finish_kernel.
#include "tc_ptx.cuh"

namespace hpfw_b200 {

constexpr int CV_CTX = HPFW_CONTEXT;           // 20 lags
constexpr int CV_LAGS = 4;                     // lags per CTA
constexpr int CV_CHUNK = 48;                   // rows (K) per pipeline stage = 6 MMA K-steps
constexpr int CV_ROWS = 72;                    // CV_CHUNK + 19 halo rows, rounded up to a multiple of 8
constexpr int CV_STAGES = 2;
constexpr uint32_t CV_ATOM_BYTES = CV_ROWS * 128;                 // 9,216: one 32-band swizzle atom of a block
constexpr uint32_t CV_PART_BYTES = 4 * CV_ATOM_BYTES;             // 36,864: hi or lo block
constexpr uint32_t CV_STAGE_BYTES = 2 * CV_PART_BYTES;            // hi + lo
// instruction descriptor: c = F32, a = b = TF32, A and B MN-major (bits 15, 16), N = 128, M = 128
constexpr uint32_t CV_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

// MN-major tf32 operands exist in one shared-memory layout only, SWIZZLE_128B with 32-byte atomicity (layout type 1): rows
// (K) of 128 bytes = 32 MN elements, the 32-byte chunk index XORed with (row & 3); 4 K-rows per swizzle atom, so the two
// halves of a K = 8 step are SBO = 512 bytes apart, and the next 32 MN elements one block atom further (LBO). The TMA mode
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B writes exactly this.
__device__ __forceinline__ uint64_t cv_desc(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(CV_ATOM_BYTES >> 4) << 16) | (32ull << 32) | (1ull << 46) | (1ull << 61);
}

// part[(split * 20 + d) * 128 * 128 + b * 128 + b'] = sum over this split's rows of Sc[u][b] * Sc[u + d][b']
__global__ void __launch_bounds__(256, 1)
cov_tc_kernel(const __grid_constant__ CUtensorMap tmHi, const __grid_constant__ CUtensorMap tmLo, int ksteps_total,
              int ksteps_per_split, float *__restrict__ part) {
    extern __shared__ uint8_t csm_raw[];
    uint8_t *csm = csm_raw + ((1024u - (smem_u32(csm_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t bar_full[CV_STAGES], bar_empty[CV_STAGES], bar_acc;
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int split = blockIdx.x, d0 = blockIdx.y * CV_LAGS;
    const int ks_begin = split * ksteps_per_split, ks_end = min(ksteps_total, ks_begin + ksteps_per_split);
    const int nsteps = max(0, ks_end - ks_begin);
    const int steps_per_chunk = CV_CHUNK / 8;
    const int nchunks = (nsteps + steps_per_chunk - 1) / steps_per_chunk;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 32) {
        for (int s = 0; s < CV_STAGES; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], CV_LAGS);      // every lag's issuing thread has read the stage
        }
        mbar_init(&bar_acc, CV_LAGS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer: block = rows [8 * (ks_begin + chunk * 6), + CV_ROWS) of the hi and the lo matrix =====
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % CV_STAGES;
            if (c >= CV_STAGES) mbar_wait(&bar_empty[s], ((c / CV_STAGES) - 1) & 1);
            const int row0 = 8 * (ks_begin + c * steps_per_chunk);
            mbar_expect_tx(&bar_full[s], CV_STAGE_BYTES);
            uint8_t *dst = csm + s * CV_STAGE_BYTES;
            for (int kb = 0; kb < 4; ++kb) {
                tma_load_2d(&tmHi, &bar_full[s], dst + kb * CV_ATOM_BYTES, kb * 32, row0);
                tma_load_2d(&tmLo, &bar_full[s], dst + CV_PART_BYTES + kb * CV_ATOM_BYTES, kb * 32, row0);
            }
        }
    } else if (warp >= 4 && lane == 0) {
        // ===== MMA issuers: one thread per lag (warps 4-7), each with its own 128-column accumulator. A 64-clk MMA leaves a
        // single issuing thread too few instructions for twelve MMAs per step (the matcher's lesson, match_tc.cu); with one
        // thread per lag each issues three per step. The descriptors of a step differ from the previous step's by a constant. =====
        const int dl = warp - 4, d = d0 + dl;
        const bool active = d < CV_CTX;
        const uint32_t dcol = tmem_base + (uint32_t)dl * 128u;
        uint32_t acc = 0;
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % CV_STAGES;
            mbar_wait(&bar_full[s], (c / CV_STAGES) & 1);
            tc_fence_after();
            if (active) {
                const uint32_t hi = smem_u32(csm + s * CV_STAGE_BYTES), lo = hi + CV_PART_BYTES;
                const int steps = min(steps_per_chunk, nsteps - c * steps_per_chunk);
                // operand A: the block at row 8k; operand B: the same block d rows further down
                uint64_t a_hi = cv_desc(hi), a_lo = cv_desc(lo), b_hi = cv_desc(hi + (uint32_t)d * 128u),
                         b_lo = cv_desc(lo + (uint32_t)d * 128u);
                for (int k = 0; k < steps; ++k) {
                    tc_mma_tf32(dcol, a_hi, b_hi, CV_IDESC, acc);
                    tc_mma_tf32(dcol, a_hi, b_lo, CV_IDESC, 1);
                    tc_mma_tf32(dcol, a_lo, b_hi, CV_IDESC, 1);
                    acc = 1;
                    a_hi += 64;     // 1,024 bytes = 8 rows further, in descriptor units of 16 bytes
                    a_lo += 64;
                    b_hi += 64;
                    b_lo += 64;
                }
            }
            tc_commit(&bar_empty[s]);
        }
        tc_commit(&bar_acc);
    }

    // ===== epilogue: warp w owns TMEM lanes 32w .. 32w+31 = rows b of the four 128 x 128 accumulators =====
    if (warp < 4) {
    mbar_wait(&bar_acc, 0);
    tc_fence_after();
    const int b = warp * 32 + lane;
    for (int dl = 0; dl < CV_LAGS; ++dl) {
        const int d = d0 + dl;
        if (d >= CV_CTX) break;
        float *dst = part + ((size_t)split * CV_CTX + d) * (128 * 128) + (size_t)b * 128;
        for (int h = 0; h < 2; ++h) {
            uint32_t v[64];
            if (nsteps > 0) {
                tmem_ld_x64(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(dl * 128 + h * 64), v);
            } else {
#pragma unroll
                for (int i = 0; i < 64; ++i) v[i] = 0u;
            }
#pragma unroll
            for (int i = 0; i < 64; i += 4)
                *reinterpret_cast<uint4 *>(dst + h * 64 + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
    }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// hi / lo: [cols][128] tf32 parts of the centred spectrogram (center_kernel, learn.cu). Writes part[splits][20][128][128] for the
// rows u < 8 * floor(nf / 8); returns that row count in *rows_done.
int cov_tc_run(hpfw_ctx *ctx, const float *d_hi, const float *d_lo, int cols, int nf, int splits, float *d_part,
               int *rows_done, cudaStream_t stream) {
    const int ksteps = nf / 8;
    *rows_done = ksteps * 8;
    const int per = (ksteps + splits - 1) / splits;
    CUtensorMap tmHi, tmLo;
    const uint64_t map_rows = (uint64_t)std::max(cols, CV_ROWS);
    HPFW_TRY(tc_make_map_2d(&tmHi, d_hi, map_rows, CV_ROWS, true));
    HPFW_TRY(tc_make_map_2d(&tmLo, d_lo, map_rows, CV_ROWS, true));
    const size_t smem = (size_t)CV_STAGES * CV_STAGE_BYTES + 1024;
    HPFW_CUDA_TRY(cudaFuncSetAttribute(cov_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    KernelScope ks(ctx, HPFW_K_OTHER, stream);
    cov_tc_kernel<<<dim3(splits, CV_CTX / CV_LAGS), 256, smem, stream>>>(tmHi, tmLo, ksteps, per, d_part);
    HPFW_CUDA_TRY(cudaGetLastError());
    return HPFW_OK;
}

}  // namespace hpfw_b200
