// hpfw_b200/csrc/project_tc.cu — stages 2+3 on the 5th-generation tensor cores (tcgen05, TMEM accumulator, TMA-fed).
//
// Same computation as project.cu (HashprintHandle::calc_frames + `filters * frames` + calc_fingerprint +
// fingerprint_to_hashprint, /root/reference/include/hpfw/core/hashprint_handle.h:79-142, parallel_collector.h:57):
//
//   hp[t] bit (63-f) = [ sum_{c<20} sum_{b<121} F[f, b*20+c] * (S[b,t+c] - S[b,t+80+c]) >= 0 ]
//
// as an implicit GEMM: per 128-frame tile D[128 x 64] = sum_c A_c[128 x 128] * B_c[64 x 128]^T with
//   A_c[t, b] = Dd[t0+t+c][b]     Dd = tf32(S[t] - S[t+80]), bands padded 121 -> 128 (a pre-pass writes it once, 512 B/row)
//   B_c[f, b] = tf32(F[f, b*20+c]) (permuted once in hpfw_set_filters)
// K = 20 taps x 128 bands, split in 32-band (128-byte) swizzle atoms: 80 x 4 tcgen05.mma.kind::tf32 (M128 N64 K8) per tile.
//
// The context window makes the 20 A_c tiles overlapping row windows of ONE [147 x 128] block, so the block is loaded once by
// TMA (4 boxes of 152 rows x 128 B, SWIZZLE_128B) and tap c is addressed by advancing the shared-memory descriptor's start
// address by c rows (c * 128 B): TMA and the MMA both derive the 128-byte swizzle phase from the shared-memory address
// bits, so a row-offset start stays consistent with how the block was written. (impl 2 = conservative variant that
// reloads the [128 x 128] window per tap; both are tested against the CUDA-core kernel and the oracle.)
// impl 3 (default) = impl 1 with fp16 operands (kind::f16): fp16 has the 10 mantissa bits of tf32, the dB differences
// (|x| <= 80) and the filter coefficients (|x| <= 1) are far inside its range, and a 128-byte swizzle atom then holds 64
// bands instead of 32: half the MMAs (160 per tile, K = 16 each), half the shared-memory operand bytes per frame — the
// kernel is bound by shared-memory operand reads (N = 64: 6 KB per 32-clk MMA) — and half the pre-pass bytes.
// B_c streams through a 3-stage TMA/mbarrier ring. The epilogue reads the TMEM accumulator with tcgen05.ld (one frame per
// thread, 64 filters in registers), thresholds and packs the 64-bit word.
#include "tc_ptx.cuh"

#include <cuda_fp16.h>

#include <cstring>

namespace hpfw_b200 {

constexpr int TC_BINS = HPFW_BINS;       // 121
constexpr int TC_BPAD = 128;             // bands padded
constexpr int TC_CTX = HPFW_CONTEXT;     // 20
constexpr int TC_LAG = HPFW_LAG;         // 80
constexpr int TC_NF = HPFW_NFILTERS;     // 64
constexpr int TC_M = 128;                // frames per tile
constexpr int TC_AROWS = 152;            // 128 + 19 rounded up to a multiple of 8
constexpr int TC_KB = 4;                 // 32-band swizzle atoms per tap
constexpr int TC_STAGES_1 = 4;          // B ring depth, impl 1: one 8 KB swizzle atom (one tap x 32 bands) per stage, so that
                                        // A (76 KB) + ring (32 KB) lets TWO CTAs share an SM: one tile's prologue and
                                        // epilogue overlap the other's MMA stream
constexpr int TC_STAGES_2 = 2;          // A+B ring depth, impl 2 (96 KB per stage)
constexpr uint32_t TC_A_ATOM_BYTES = TC_AROWS * 128;        // 19,456
constexpr uint32_t TC_A1_ATOM_BYTES = TC_M * 128;           // 16,384 (impl 2)
constexpr uint32_t TC_B_ATOM_BYTES = TC_NF * 128;           // 8,192
constexpr uint32_t TC_B_STAGE_BYTES = TC_KB * TC_B_ATOM_BYTES;   // 32,768
constexpr uint32_t TC_TMEM_COLS = 64;
// instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6), a=b=TF32 [7,10)/[10,13), K-major both, N>>3 [17,23), M>>4 [24,29)
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((TC_NF >> 3) << 17) | ((TC_M >> 4) << 24);

struct TcTile {
    int32_t track;
    int32_t t0;
};
struct TcTrack {
    int64_t row_base;    // first row of this track in the concatenated Dd matrix
    int64_t out_start;
    int32_t n_out;       // hashprint words
    int32_t pad;
};

// ---- pre-pass: Dd[row][b] = tf32(S[t][b] - S[t+80][b]), bands 121..127 = 0 ------------------------------------------------
template <int HALF>   // 0: tf32 values in float storage; 1: fp16
__global__ void __launch_bounds__(256)
tc_delta_kernel(const float *__restrict__ spectro, const int64_t *__restrict__ col_start, const int64_t *__restrict__ row_base,
                const int32_t *__restrict__ rows, int n_tracks, void *__restrict__ dd_) {
    const int trk = blockIdx.y;
    const int nr = rows[trk];
    const float *S = spectro + col_start[trk] * TC_BINS;
    float *D = static_cast<float *>(dd_) + row_base[trk] * TC_BPAD;
    __half *Dh = static_cast<__half *>(dd_) + row_base[trk] * TC_BPAD;
    const int64_t total = (int64_t)nr * TC_BPAD;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = i >> 7;
        const int b = (int)(i & 127);
        float v = 0.f;
        if (b < TC_BINS) v = S[t * TC_BINS + b] - S[(t + TC_LAG) * TC_BINS + b];
        if (HALF) {
            Dh[i] = __float2half_rn(v);
        } else {
            uint32_t r;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
            D[i] = __uint_as_float(r);
        }
    }
}

// ---- main kernel ------------------------------------------------------------------------------------------------------------
// instruction descriptor for fp16 operands: c=F32, a=b=F16 (0), K-major both
constexpr uint32_t TC_IDESC_H = (1u << 4) | ((TC_NF >> 3) << 17) | ((TC_M >> 4) << 24);
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}

template <int IMPL>   // 1: one A block + row-offset descriptors; 2: A window reloaded per tap; 3: as 1 with fp16 operands
__global__ void __launch_bounds__(128, IMPL == 1 ? 2 : (IMPL == 3 ? 3 : 1))
project_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const TcTile *__restrict__ tiles, const TcTrack *__restrict__ tracks, uint64_t *__restrict__ hp_out) {
    extern __shared__ uint8_t tsm_raw[];
    // SWIZZLE_128B atoms must sit on 1024-byte boundaries: align the dynamic region by hand (1 KB of slack is allocated)
    uint8_t *tsm = tsm_raw + ((1024u - (smem_u32(tsm_raw) & 1023u)) & 1023u);
    // carve: [A region][B stages], every atom a multiple of 1024 bytes
    constexpr bool ONE = IMPL != 2;                 // one A block per tile
    constexpr int KB = (IMPL == 3) ? 2 : TC_KB;     // 128-byte swizzle atoms per tap: 64 fp16 or 32 tf32 bands each
    constexpr int XS = (IMPL == 3) ? 64 : 32;       // bands per atom (tensor-map x step)
    constexpr int TC_STAGES = ONE ? TC_STAGES_1 : TC_STAGES_2;
    constexpr uint32_t A_BYTES = ONE ? KB * TC_A_ATOM_BYTES : TC_STAGES * TC_KB * TC_A1_ATOM_BYTES;
    uint8_t *smA = tsm;
    uint8_t *smB = tsm + A_BYTES;
    __shared__ __align__(8) uint64_t bar_a, bar_full[TC_STAGES], bar_empty[TC_STAGES], bar_acc;
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const TcTile tile = tiles[blockIdx.x];
    const TcTrack trk = tracks[tile.track];
    const int row0 = (int)(trk.row_base + tile.t0);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 32) {
        mbar_init(&bar_a, 1);
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], 1);
        }
        mbar_init(&bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer =====
        if (ONE) {
            mbar_expect_tx(&bar_a, KB * TC_A_ATOM_BYTES);
            for (int kb = 0; kb < KB; ++kb) tma_load_2d(&tmA, &bar_a, smA + kb * TC_A_ATOM_BYTES, kb * XS, row0);
        }
        if (ONE) {
            for (int it = 0; it < TC_CTX * KB; ++it) {        // stage = one (tap, band atom) of B
                const int s = it % TC_STAGES, c = it / KB, kb = it % KB;
                if (it >= TC_STAGES) mbar_wait(&bar_empty[s], ((it / TC_STAGES) - 1) & 1);
                mbar_expect_tx(&bar_full[s], TC_B_ATOM_BYTES);
                tma_load_2d(&tmB, &bar_full[s], smB + s * TC_B_ATOM_BYTES, kb * XS, c * TC_NF);
            }
        } else {
            for (int c = 0; c < TC_CTX; ++c) {
                const int s = c % TC_STAGES;
                if (c >= TC_STAGES) mbar_wait(&bar_empty[s], ((c / TC_STAGES) - 1) & 1);
                mbar_expect_tx(&bar_full[s], TC_B_STAGE_BYTES + TC_KB * TC_A1_ATOM_BYTES);
                for (int kb = 0; kb < TC_KB; ++kb)
                    tma_load_2d(&tmB, &bar_full[s], smB + s * TC_B_STAGE_BYTES + kb * TC_B_ATOM_BYTES, kb * 32, c * TC_NF);
                for (int kb = 0; kb < TC_KB; ++kb)
                    tma_load_2d(&tmA, &bar_full[s], smA + (s * TC_KB + kb) * TC_A1_ATOM_BYTES, kb * 32, row0 + c);
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ===== MMA issuer =====
        if (ONE) mbar_wait(&bar_a, 0);
        uint32_t acc = 0;
        if (ONE) {
            for (int it = 0; it < TC_CTX * KB; ++it) {
                const int s = it % TC_STAGES, c = it / KB, kb = it % KB;
                mbar_wait(&bar_full[s], (it / TC_STAGES) & 1);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smA + kb * TC_A_ATOM_BYTES) + c * 128;
                const uint32_t b_addr = smem_u32(smB + s * TC_B_ATOM_BYTES);
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {     // 8 tf32 or 16 fp16 = 32 bytes per MMA along K inside the 128-byte atom
                    if (IMPL == 3)
                        tc_mma_f16(tmem_base, smem_desc_sw128(a_addr + k4 * 32), smem_desc_sw128(b_addr + k4 * 32), TC_IDESC_H, acc);
                    else
                        tc_mma_tf32(tmem_base, smem_desc_sw128(a_addr + k4 * 32), smem_desc_sw128(b_addr + k4 * 32), TC_IDESC, acc);
                    acc = 1;
                }
                tc_commit(&bar_empty[s]);     // arrives when the MMAs above have finished reading this stage
            }
        } else {
            for (int c = 0; c < TC_CTX; ++c) {
                const int s = c % TC_STAGES;
                mbar_wait(&bar_full[s], (c / TC_STAGES) & 1);
                tc_fence_after();
#pragma unroll
                for (int kb = 0; kb < TC_KB; ++kb) {
                    const uint32_t a_addr = smem_u32(smA + (s * TC_KB + kb) * TC_A1_ATOM_BYTES);
                    const uint32_t b_addr = smem_u32(smB + s * TC_B_STAGE_BYTES + kb * TC_B_ATOM_BYTES);
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        tc_mma_tf32(tmem_base, smem_desc_sw128(a_addr + k4 * 32), smem_desc_sw128(b_addr + k4 * 32), TC_IDESC, acc);
                        acc = 1;
                    }
                }
                tc_commit(&bar_empty[s]);
            }
        }
        tc_commit(&bar_acc);              // accumulator complete
    }

    // ===== epilogue: all four warps; warp w owns TMEM lanes (= frames) 32w .. 32w+31 =====
    mbar_wait(&bar_acc, 0);
    tc_fence_after();
    uint32_t v[64];
    tmem_ld_x64(tmem_base + ((uint32_t)(warp * 32) << 16), v);
    uint64_t word = 0;
#pragma unroll
    for (int f = 0; f < 64; ++f) word |= (uint64_t)(__uint_as_float(v[f]) >= 0.f ? 1u : 0u) << (63 - f);
    const int t = tile.t0 + warp * 32 + lane;
    if (t < trk.n_out) hp_out[trk.out_start + t] = word;

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
    }
}

// ---- impl 4 / 5: NT tiles per CTA share the filter stream -----------------------------------------------------------------
// In project_tc_kernel every 128-frame tile streams all 20 taps of the fp16 filters (320 KB) from L2 for its 160 MMAs: at the
// full tensor rate that is 64 B/clk per SM, 9.5 KB/clk over the chip, more than the L2 -> SM path delivers (~6.3 KB/clk), and
// one thread issues every MMA of the CTA (a 32-clk MMA leaves it ~30 instructions). Here a CTA owns NT consecutive tiles: one
// A block, one 64-column TMEM accumulator and ONE ISSUING THREAD per tile, all fed from the same ring of filter stages (a stage
// is released when every tile's issuer has committed it), so the filter stream per frame drops by NT and the issue budget per
// thread grows by NT. Per tile the MMAs, their order and their operands are those of impl 3: identical hashprints.
// Warps: 0-3 epilogue (warp 0 lane 0 also the TMA producer), 4 .. 4+NT-1 the issuers.
template <int NT>
__global__ void __launch_bounds__(128 + 32 * NT, NT == 2 ? 2 : 1)
project_tc_multi_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const TcTile *__restrict__ tiles, int n_tiles, const TcTrack *__restrict__ tracks,
                        uint64_t *__restrict__ hp_out) {
    extern __shared__ uint8_t msm_raw[];
    uint8_t *msm = msm_raw + ((1024u - (smem_u32(msm_raw) & 1023u)) & 1023u);
    constexpr int KB = 2, XS = 64, STAGES = TC_STAGES_1;
    constexpr uint32_t A_BLOCK = KB * TC_A_ATOM_BYTES;          // 38,912 bytes per tile
    uint8_t *smA = msm;
    uint8_t *smB = msm + NT * A_BLOCK;
    __shared__ __align__(8) uint64_t bar_a[NT], bar_full[STAGES], bar_empty[STAGES], bar_acc[NT];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int first = blockIdx.x * NT;
    const int nt = min(NT, n_tiles - first);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"((uint32_t)(NT * TC_TMEM_COLS)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 32) {
        for (int i = 0; i < NT; ++i) {
            mbar_init(&bar_a[i], 1);
            mbar_init(&bar_acc[i], 1);
        }
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], (uint32_t)nt);      // every tile's issuer has read the stage
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer: the A block of every tile, then the 40 filter stages =====
        for (int i = 0; i < nt; ++i) {
            const TcTile tile = tiles[first + i];
            const int row0 = (int)(tracks[tile.track].row_base + tile.t0);
            mbar_expect_tx(&bar_a[i], A_BLOCK);
            for (int kb = 0; kb < KB; ++kb) tma_load_2d(&tmA, &bar_a[i], smA + i * A_BLOCK + kb * TC_A_ATOM_BYTES, kb * XS, row0);
        }
        for (int it = 0; it < TC_CTX * KB; ++it) {
            const int s = it % STAGES, c = it / KB, kb = it % KB;
            if (it >= STAGES) mbar_wait(&bar_empty[s], ((it / STAGES) - 1) & 1);
            mbar_expect_tx(&bar_full[s], TC_B_ATOM_BYTES);
            tma_load_2d(&tmB, &bar_full[s], smB + s * TC_B_ATOM_BYTES, kb * XS, c * TC_NF);
        }
    } else if (warp >= 4 && warp - 4 < nt && lane == 0) {
        // ===== MMA issuer of tile (warp - 4): descriptors advance by constants, two 64-bit adds per MMA =====
        const int i = warp - 4;
        mbar_wait(&bar_a[i], 0);
        tc_fence_after();
        const uint64_t a_tile = smem_desc_sw128(smem_u32(smA + i * A_BLOCK));
        const uint64_t b_ring = smem_desc_sw128(smem_u32(smB));
        const uint32_t d = tmem_base + (uint32_t)(i * TC_TMEM_COLS);
        uint32_t acc = 0;
        for (int it = 0; it < TC_CTX * KB; ++it) {
            const int s = it % STAGES, c = it / KB, kb = it % KB;
            mbar_wait(&bar_full[s], (it / STAGES) & 1);
            tc_fence_after();
            const uint64_t a_desc = a_tile + (uint64_t)((kb * TC_A_ATOM_BYTES + c * 128) >> 4);
            const uint64_t b_desc = b_ring + (uint64_t)((s * TC_B_ATOM_BYTES) >> 4);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {       // 16 fp16 = 32 bytes per MMA along K inside the 128-byte swizzle atom
                tc_mma_f16(d, a_desc + (uint64_t)(k4 * 2), b_desc + (uint64_t)(k4 * 2), TC_IDESC_H, acc);
                acc = 1;
            }
            tc_commit(&bar_empty[s]);
        }
        tc_commit(&bar_acc[i]);
    }

    // ===== epilogue: warps 0-3; warp w owns TMEM lanes (= frames) 32w .. 32w+31 of every tile's accumulator =====
    if (warp < 4) {
        for (int i = 0; i < nt; ++i) {
            const TcTile tile = tiles[first + i];
            const TcTrack trk = tracks[tile.track];
            mbar_wait(&bar_acc[i], 0);
            tc_fence_after();
            uint32_t v[64];
            tmem_ld_x64(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(i * TC_TMEM_COLS), v);
            uint64_t word = 0;
#pragma unroll
            for (int f = 0; f < 64; ++f) word |= (uint64_t)(__uint_as_float(v[f]) >= 0.f ? 1u : 0u) << (63 - f);
            const int t = tile.t0 + warp * 32 + lane;
            if (t < trk.n_out) hp_out[trk.out_start + t] = word;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(NT * TC_TMEM_COLS))
                     : "memory");
    }
}

static int tc_make_map_2d_t(CUtensorMap *map, const void *base, uint64_t rows, uint32_t box_rows, bool atom32, bool half);
int tc_make_map_2d(CUtensorMap *map, const void *base, uint64_t rows, uint32_t box_rows, bool atom32) {
    return tc_make_map_2d_t(map, base, rows, box_rows, atom32, false);
}
static int tc_make_map_2d_t(CUtensorMap *map, const void *base, uint64_t rows, uint32_t box_rows, bool atom32, bool half) {
    // row-major [rows][128] float (pitch 512 B, box = 32 bands x box_rows rows) or fp16 (pitch 256 B, box = 64 bands):
    // dim0 = 128 bands (contiguous), dim1 = rows; a box row is one 128-byte swizzle row either way
    const cuuint64_t dims[2] = {TC_BPAD, rows};
    const cuuint64_t strides[1] = {TC_BPAD * (half ? sizeof(__half) : sizeof(float))};
    const cuuint32_t box[2] = {half ? 64u : 32u, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    // the driver entry point is resolved through the runtime so that the library does not link libcuda.so (it must load,
    // and fail loudly, on machines without a driver)
    typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeTiled encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        HPFW_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn || q != cudaDriverEntryPointSuccess) HPFW_FAIL(HPFW_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
        encode = reinterpret_cast<EncodeTiled>(fn);
    }
    CUresult r = encode(map, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) HPFW_FAIL(HPFW_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return HPFW_OK;
}

// filters (column-major 64 x 2420, row index b*20+c) -> tf32 [c][f][b padded to 128]
int project_tc_set_filters(hpfw_ctx *ctx, const float *f) {
    std::vector<float> perm((size_t)TC_CTX * TC_NF * TC_BPAD, 0.f);
    for (int c = 0; c < TC_CTX; ++c)
        for (int fi = 0; fi < TC_NF; ++fi)
            for (int b = 0; b < TC_BINS; ++b) {
                float v = f[fi + (size_t)TC_NF * (b * TC_CTX + c)];
                uint32_t u;
                memcpy(&u, &v, 4);
                // round to nearest, ties away (cvt.rna.tf32): keep 10 mantissa bits
                if ((u & 0x7F800000u) != 0x7F800000u) u = (u + 0x1000u) & 0xFFFFE000u;
                memcpy(&v, &u, 4);
                perm[((size_t)c * TC_NF + fi) * TC_BPAD + b] = v;
            }
    HPFW_TRY(ctx->filters_tc.reserve(sizeof(float) * perm.size()));
    HPFW_CUDA_TRY(cudaMemcpy(ctx->filters_tc.ptr, perm.data(), sizeof(float) * perm.size(), cudaMemcpyHostToDevice));
    // fp16 copy for impl 3, same [c][f][b] order (round to nearest even from the original floats)
    std::vector<__half> permh(perm.size());
    for (int c = 0; c < TC_CTX; ++c)
        for (int fi = 0; fi < TC_NF; ++fi)
            for (int b = 0; b < TC_BPAD; ++b)
                permh[((size_t)c * TC_NF + fi) * TC_BPAD + b] =
                    __float2half_rn(b < TC_BINS ? f[fi + (size_t)TC_NF * (b * TC_CTX + c)] : 0.f);
    HPFW_TRY(ctx->filters_tc16.reserve(sizeof(__half) * permh.size()));
    HPFW_CUDA_TRY(cudaMemcpy(ctx->filters_tc16.ptr, permh.data(), sizeof(__half) * permh.size(), cudaMemcpyHostToDevice));
    return HPFW_OK;
}

int project_tc_run(hpfw_ctx *ctx, int impl, const float *d_spectro, const int64_t *col_offsets, int n, uint64_t *d_hp,
                   cudaStream_t stream) {
    std::vector<TcTrack> tracks((size_t)std::max(n, 1));
    std::vector<TcTile> tiles;
    std::vector<int64_t> col_start((size_t)n), row_base((size_t)n);
    std::vector<int32_t> rows((size_t)n);
    int64_t out = 0, rb = 0;
    int max_rows = 0;
    for (int i = 0; i < n; ++i) {
        const int64_t cols = col_offsets[i + 1] - col_offsets[i];
        if (cols < 0 || cols > (int64_t(1) << 30)) HPFW_FAIL(HPFW_ERR_ARG, "bad column count for spectrogram %d", i);
        const int64_t n_out = std::max<int64_t>(0, cols - (TC_CTX - 1) - TC_LAG);
        const int64_t nr = n_out > 0 ? cols - TC_LAG : 0;       // rows of Dd this track needs: n_out + 19
        col_start[i] = col_offsets[i];
        row_base[i] = rb;
        rows[i] = (int32_t)nr;
        tracks[i] = {rb, out, (int32_t)n_out, 0};
        for (int64_t t0 = 0; t0 < n_out; t0 += TC_M) tiles.push_back({i, (int32_t)t0});
        out += n_out;
        rb += nr;
        max_rows = std::max<int>(max_rows, (int)nr);
    }
    if (tiles.empty()) return HPFW_OK;
    // device metadata
    const size_t b_tracks = sizeof(TcTrack) * tracks.size(), b_tiles = sizeof(TcTile) * tiles.size();
    const size_t b_i64 = sizeof(int64_t) * (size_t)n, b_i32 = sizeof(int32_t) * (size_t)n;
    const size_t o_tiles = b_tracks, o_cs = o_tiles + ((b_tiles + 7) & ~size_t(7)), o_rb = o_cs + b_i64, o_rows = o_rb + b_i64;
    const size_t total = o_rows + b_i32;
    HPFW_CUDA_TRY(cudaEventSynchronize(ctx->pin_in_free));
    HPFW_TRY(ctx->pin_in.reserve(total));
    HPFW_TRY(ctx->colmeta.reserve(total));
    char *h = ctx->pin_in.as<char>();
    memcpy(h, tracks.data(), b_tracks);
    memcpy(h + o_tiles, tiles.data(), b_tiles);
    memcpy(h + o_cs, col_start.data(), b_i64);
    memcpy(h + o_rb, row_base.data(), b_i64);
    memcpy(h + o_rows, rows.data(), b_i32);
    HPFW_CUDA_TRY(cudaMemcpyAsync(ctx->colmeta.ptr, h, total, cudaMemcpyHostToDevice, stream));
    HPFW_CUDA_TRY(cudaEventRecord(ctx->pin_in_free, stream));
    const char *dm = ctx->colmeta.as<char>();
    // at least one TMA box of rows (rows past the last track are never part of a stored frame)
    const uint64_t map_rows = std::max<uint64_t>((uint64_t)rb, 256);
    HPFW_TRY(ctx->delta_tc.reserve(sizeof(float) * (size_t)map_rows * TC_BPAD));
    const bool half = impl >= 3;
    {
        KernelScope ks(ctx, HPFW_K_PROJECT, stream);
        const int gx = std::max(1, std::min(64, (max_rows * TC_BPAD + 256 * 8 - 1) / (256 * 8)));
        if (half)
            tc_delta_kernel<1><<<dim3(gx, n), 256, 0, stream>>>(d_spectro, reinterpret_cast<const int64_t *>(dm + o_cs),
                                                                reinterpret_cast<const int64_t *>(dm + o_rb),
                                                                reinterpret_cast<const int32_t *>(dm + o_rows), n,
                                                                ctx->delta_tc.ptr);
        else
            tc_delta_kernel<0><<<dim3(gx, n), 256, 0, stream>>>(d_spectro, reinterpret_cast<const int64_t *>(dm + o_cs),
                                                                reinterpret_cast<const int64_t *>(dm + o_rb),
                                                                reinterpret_cast<const int32_t *>(dm + o_rows), n,
                                                                ctx->delta_tc.ptr);
    }
    CUtensorMap tmA, tmB;
    HPFW_TRY(tc_make_map_2d_t(&tmA, ctx->delta_tc.ptr, map_rows, impl != 2 ? TC_AROWS : TC_M, false, half));
    HPFW_TRY(tc_make_map_2d_t(&tmB, half ? ctx->filters_tc16.ptr : ctx->filters_tc.ptr, (uint64_t)TC_CTX * TC_NF, TC_NF, false,
                              half));
    const size_t smem1 = TC_KB * TC_A_ATOM_BYTES + TC_STAGES_1 * TC_B_ATOM_BYTES + 1024;
    const size_t smem2 = TC_STAGES_2 * TC_KB * TC_A1_ATOM_BYTES + TC_STAGES_2 * TC_B_STAGE_BYTES + 1024;
    const size_t smem3 = 2 * TC_A_ATOM_BYTES + TC_STAGES_1 * TC_B_ATOM_BYTES + 1024;
    {
        KernelScope ks(ctx, HPFW_K_PROJECT, stream);
        if (impl == 4 || impl == 5) {
            const int nt = impl == 4 ? 2 : 4, n_tiles = (int)tiles.size();
            const size_t smem_m = (size_t)nt * 2 * TC_A_ATOM_BYTES + TC_STAGES_1 * TC_B_ATOM_BYTES + 1024;
            if (nt == 2) {
                HPFW_CUDA_TRY(cudaFuncSetAttribute(project_tc_multi_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_m));
                project_tc_multi_kernel<2><<<(unsigned)((n_tiles + 1) / 2), 192, smem_m, stream>>>(
                    tmA, tmB, reinterpret_cast<const TcTile *>(dm + o_tiles), n_tiles, reinterpret_cast<const TcTrack *>(dm), d_hp);
            } else {
                HPFW_CUDA_TRY(cudaFuncSetAttribute(project_tc_multi_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_m));
                project_tc_multi_kernel<4><<<(unsigned)((n_tiles + 3) / 4), 256, smem_m, stream>>>(
                    tmA, tmB, reinterpret_cast<const TcTile *>(dm + o_tiles), n_tiles, reinterpret_cast<const TcTrack *>(dm), d_hp);
            }
        } else if (impl == 3) {
            HPFW_CUDA_TRY(cudaFuncSetAttribute(project_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
            project_tc_kernel<3><<<(unsigned)tiles.size(), 128, smem3, stream>>>(
                tmA, tmB, reinterpret_cast<const TcTile *>(dm + o_tiles), reinterpret_cast<const TcTrack *>(dm), d_hp);
        } else if (impl == 1) {
            HPFW_CUDA_TRY(cudaFuncSetAttribute(project_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
            project_tc_kernel<1><<<(unsigned)tiles.size(), 128, smem1, stream>>>(
                tmA, tmB, reinterpret_cast<const TcTile *>(dm + o_tiles), reinterpret_cast<const TcTrack *>(dm), d_hp);
        } else {
            HPFW_CUDA_TRY(cudaFuncSetAttribute(project_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            project_tc_kernel<2><<<(unsigned)tiles.size(), 128, smem2, stream>>>(
                tmA, tmB, reinterpret_cast<const TcTile *>(dm + o_tiles), reinterpret_cast<const TcTrack *>(dm), d_hp);
        }
    }
    HPFW_CUDA_TRY(cudaGetLastError());
    return HPFW_OK;
}

}  // namespace hpfw_b200
