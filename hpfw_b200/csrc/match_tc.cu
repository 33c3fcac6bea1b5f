// hpfw_b200/csrc/match_tc.cu — stage 4 on the 5th-generation tensor cores: the Hamming cross-correlation of
// db::MemoryStorage::find (/root/reference/include/hpfw/audioproblems/live-song-id/storage.h:27-64) as an EXACT GEMM.
//
// For a query q[0..k) and a reference track r[0..n):  D[i] = sum_j popc(q[j] ^ r[i+j]). With every bit b of a word mapped to
// s(b) = +1 (set) / -1 (clear),  sum_bits s(q) * s(r) = 64 - 2 * popc(q ^ r),  so
//     D[i] = (64 k - dot[i]) / 2,      dot[i] = sum_{j < k} sum_{b < 64} s(q[j])_b * s(r[i + j])_b
// and dot[] is a GEMM whose "offset" operand is a HANKEL matrix: row i of it is the expansion of r[i .. i+k), i.e. row i+1
// is row i moved one word along K. Every product is +-1 and |dot| <= 64 * 4096 = 2^18, so the sums are exact in the s32
// accumulators of tcgen05.mma.kind::i8 (signed bytes) and in the f32 accumulators of kind::mxf4.block_scale (e2m1 nibbles,
// unit scales; integers below 2^24): distances and rankings stay bit-exact. The two encodings are the template parameter.
//
// Mapping (one CTA per SM, persistent over (query group, tile) items):
//   * M = 128 queries of one group (TMEM lanes), N = the alignment offsets of one tile of one track (2 MMAs of N = 256 or
//     240 = the TMEM columns), K = 64 k, walked one 64-bit word at a time (2 MMAs of K = 32 bytes, or 1 of K = 64 nibbles).
//   * offset operand: the tile's reference words are expanded ONCE per K chunk into shared memory as
//     R[16-byte chunk c][row w][16 B] — the no-swizzle K-major canonical layout with the 8-row core matrices of one chunk
//     column contiguous (SBO = 128 B, LBO = rows * 16 B). Row w+1 is 16 bytes after row w, so the operand for query word j
//     is the SAME buffer addressed through a descriptor whose start address is moved by j rows: no shifted copy is made,
//     a tile reads (tile + k) words instead of tile * k. Rows beyond the end of the track are zeros and contribute
//     nothing, which is exactly the reference's truncation of a query that is longer than the track (storage.h:34-38).
//   * query operand: expanded once per call by xt_expand_queries_kernel to [word j][chunk c][query m][16 B] (the words
//     beyond a shorter query's end are zeros), streamed by 1-D bulk TMA copies through an mbarrier ring (fp4: ten 16 KB
//     stages of 4 words, int8: four 32 KB stages; the 1.6 / 3 MB of a group stay in L2 for all 148 CTAs).
//   * warp roles: 0 = TMA producer, 1 = MMA issuer of the first N half, 2-5 = reference expanders (double-buffered R),
//     6-13 = epilogue, 14 = MMA issuer of the second N half (LAG stages behind the first).
//   * epilogue: tcgen05.ld 32 / 64 columns at a time (warps 6-9: the first N half of the tile, warps 10-13: the second,
//     which the issuer runs a few stages behind the first so that the TMEM drains overlap MMAs); per query (lane)
//     the maximum of (dot, lowest column) over the valid offsets; one 64-bit atomicMin per (query, tile, column half) into
//     best[query][track], the array match_kernel and topk_kernel share — its key carries the offset, so ties resolve to the
//     lowest offset as the reference's strict '<' does.
#include "matcher.cuh"
#include "tc_ptx.cuh"

#include <algorithm>
#include <type_traits>

namespace hpfw_b200 {

// Two operand encodings share the kernel (template parameter F4):
//   F4 = 0: signed bytes +1 / -1, tcgen05.mma.kind::i8, s32 accumulators; a 64-bit word is 64 bytes = four 16-byte K chunks
//           = 2 MMAs of K = 32; tile = 512 offsets = 2 x N(256) = all 512 TMEM columns.
//   F4 = 1: e2m1 nibbles +1.0 / -1.0, tcgen05.mma.kind::mxf4.block_scale with every UE8M0 scale = 1.0, f32 accumulators
//           (sums of +-1 below 2^24 are exact in f32); a word is 32 bytes = two K chunks = 1 MMA of K = 64, i.e. twice the
//           word rate per clock; tile = 480 offsets = 2 x N(240), the last 32 TMEM columns hold the scale factors.
template <int F4>
struct Xt {
    static constexpr int NOFF = F4 ? XT_NOFF_F4 : XT_NOFF;         // offsets per tile
    static constexpr int HALF = NOFF / 2;                          // N of one MMA
    static constexpr int CPW = F4 ? 2 : 4;                         // 16-byte K chunks per word
    static constexpr int JS = F4 ? XT_JS_F4 : XT_JS;               // query words per TMA stage (32 KB either way)
    static constexpr int JCMAX = F4 ? 448 : 192;                   // words per K chunk (multiple of JS)
    static constexpr int NR = NOFF + JCMAX;                        // rows of an expanded reference buffer
    static constexpr uint32_t R_LBO = NR * 16;                     // between the 16-byte K chunks of a row
    static constexpr uint32_t R_BYTES = CPW * R_LBO;               // 45,056 / 29,696
    static constexpr uint32_t Q_LBO = XT_NQ * 16;                  // 2,048
    static constexpr uint32_t QWORD_BYTES = CPW * Q_LBO;           // one query word of a group: 8,192 / 4,096
    static constexpr uint32_t STAGE_BYTES = JS * QWORD_BYTES;      // int8: 32,768; fp4: 16,384
    static constexpr int KSTEPS = F4 ? 1 : 2;                      // MMAs (per N half) per word
    static constexpr int ECH = F4 ? 32 : 64;                       // TMEM columns per epilogue load
    // The two N halves of a tile have their own accumulator barriers AND their own issuing thread (warp 1 and warp 14): the
    // second half's issuer stays LAG stages behind the first (it waits for the first half to have issued stage i + LAG, or the
    // tile's last stage), so at a tile boundary the first half finishes LAG stages early and is drained by its epilogue warps
    // while the second half's last MMAs run, and the second half is drained while the next tile's first half starts. The TMEM
    // drain (128 lanes x 240 columns at 64 B/clk = 1,920 clk per half, ~2,500 with latencies) is hidden when
    // LAG * JS * (clk per word and half) covers it. A stage stays in the ring until both halves have read it, so the ring holds
    // LAG + 1 stages in use plus the prefetch. Two issuing threads also halve the instruction budget each must meet (one MMA
    // per 240 clk instead of 120): a single thread walking both halves through a queue of stage descriptors could not keep the
    // tensor pipe fed (measured: -12 % with the queue, -25 % with 4-word stages).
#ifndef XT_F4_STAGES
#define XT_F4_STAGES 5
#endif
#ifndef XT_F4_LAG
#define XT_F4_LAG 1
#endif
    static constexpr int STAGES = F4 ? XT_F4_STAGES : 4;           // query ring depth
    static constexpr int LAG = F4 ? XT_F4_LAG : 0;                 // stages the second half trails the first
    static_assert(JCMAX % JS == 0, "chunk boundaries must be stage boundaries");
    static_assert(LAG + 2 <= STAGES, "the ring must hold the lag plus at least one stage in flight");
};
constexpr int XT_THREADS = 480;                                // TMA, MMA (first half), 4 expander, 8 epilogue warps, MMA (second half)
constexpr int XT_BIAS = (1 << 18) + 1;                         // dot + bias >= 1 for every valid offset
constexpr uint32_t XT_SF_COL = 480;                            // F4: TMEM columns 480..511 = scale factors, all 1.0

// 4 bits -> 4 bytes of +1 / -1
__device__ __forceinline__ uint32_t xt_nib(uint32_t n) {
    const uint32_t x = (n * 0x00204081u) & 0x01010101u;
    return ~(x * 0xFEu);
}
__device__ __forceinline__ uint4 xt_expand16(uint32_t h) {
    return make_uint4(xt_nib(h & 15u), xt_nib((h >> 4) & 15u), xt_nib((h >> 8) & 15u), xt_nib((h >> 12) & 15u));
}

// 8 bits -> 8 e2m1 nibbles: +1.0 = 0b0010, -1.0 = 0b1010
__device__ __forceinline__ uint32_t xt_nib8(uint32_t b) {
    uint32_t x = (b | (b << 12)) & 0x000F000Fu;
    x = (x | (x << 6)) & 0x03030303u;
    x = (x | (x << 3)) & 0x11111111u;       // bit i of b at position 4 i
    return 0xAAAAAAAAu ^ (x << 3);          // clear the sign bit where the hashprint bit is set
}
__device__ __forceinline__ uint4 xt_expand32_f4(uint32_t h) {
    return make_uint4(xt_nib8(h & 255u), xt_nib8((h >> 8) & 255u), xt_nib8((h >> 16) & 255u), xt_nib8(h >> 24));
}
// chunk c of a word in the operand encoding
template <int F4>
__device__ __forceinline__ uint4 xt_chunk(uint64_t w, int c) {
    if (F4) return xt_expand32_f4((uint32_t)(w >> (32 * c)));
    return xt_expand16((uint32_t)(w >> (16 * c)) & 0xFFFFu);
}

// shared-memory matrix descriptor, K-major, no swizzle: start >> 4 [0,14), LBO >> 4 [16,30) (between the two 16-byte K
// chunks of a K = 32 step), SBO >> 4 [32,46) (between 8-row core matrices), version 1 [46,48), layout type 0 [61,64)
__device__ __forceinline__ uint64_t xt_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor: D = S32, A = B = signed 8-bit, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t xt_idesc(int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
// block-scaled instruction descriptor: A = B = e2m1 (MXF4 format 1), K-major, UE8M0 scales (bit 23), M = 128, N = n, K = 64
__device__ __forceinline__ uint32_t xt_idesc_f4(int n) {
    return (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | (1u << 23) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void tc_mma_mxf4(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t sfa,
                                            uint32_t sfb, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%4], [%5], p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(sfa), "r"(sfb), "r"(acc) : "memory");
}
// tcgen05.ld.32x32b.x32: 32 consecutive TMEM columns of this thread's lane
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// the same value into 32 consecutive TMEM columns of this warp's 32 lanes
__device__ __forceinline__ void tmem_fill_x32(uint32_t taddr, uint32_t x) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
        "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
        ::"r"(taddr), "r"(x) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// qexp[group].exp_off + ((j * CPW + c) * 128 + m) * 16: chunk c of word j of the group's m-th query
template <int F4>
__global__ void __launch_bounds__(XT_NQ)
xt_expand_queries_kernel(const uint64_t *__restrict__ qwords, const int64_t *__restrict__ qstart,
                         const XtGroup *__restrict__ groups, const int32_t *__restrict__ row_q,
                         const int32_t *__restrict__ row_k, uint8_t *__restrict__ qexp) {
    using G = Xt<F4>;
    const int g = blockIdx.y, j = blockIdx.x, m = threadIdx.x;
    const XtGroup grp = groups[g];
    const int kpad = ((grp.kmax < 1 ? 1 : grp.kmax) + G::JS - 1) / G::JS * G::JS;
    if (j >= kpad) return;
    const int q = row_q[g * XT_NQ + m];
    const bool valid = q >= 0 && j < row_k[g * XT_NQ + m];
    const uint64_t w = valid ? qwords[qstart[q] + j] : 0ull;
    uint4 *dst = reinterpret_cast<uint4 *>(qexp + grp.exp_off) + (size_t)j * G::CPW * XT_NQ + m;
#pragma unroll
    for (int c = 0; c < G::CPW; ++c) dst[c * XT_NQ] = valid ? xt_chunk<F4>(w, c) : make_uint4(0u, 0u, 0u, 0u);
}

struct XtItem {
    int g, track, tile_start, n_r;
    int64_t tbeg;
    int kmax;    // >= 1
    int need;    // columns of this tile that hold a valid offset for at least one query of the group (<= 0: skip)
    int nchunks, jc;
};

template <int F4>
__device__ __forceinline__ XtItem xt_item(long long item, int n_tiles, const MatchTile *__restrict__ tiles,
                                          const int64_t *__restrict__ track_start, const XtGroup *__restrict__ groups) {
    using G = Xt<F4>;
    XtItem it;
    it.g = int(item / n_tiles);
    const MatchTile t = tiles[item % n_tiles];
    const XtGroup grp = groups[it.g];
    it.track = t.track;
    it.tile_start = t.start;
    it.tbeg = track_start[t.track];
    it.n_r = int(track_start[t.track + 1] - it.tbeg);
    it.kmax = max(grp.kmax, 1);
    const int last_valid = it.n_r - min(grp.kmin, it.n_r);    // the shortest query reaches the furthest offset
    it.need = min(G::NOFF, last_valid - t.start + 1);
    it.nchunks = (it.kmax + G::JCMAX - 1) / G::JCMAX;
    it.jc = ((it.kmax + it.nchunks - 1) / it.nchunks + G::JS - 1) / G::JS * G::JS;
    return it;
}

template <int F4>
__global__ void __launch_bounds__(XT_THREADS, 1)
match_tc_kernel(const uint64_t *__restrict__ words, const int64_t *__restrict__ track_start,
                const MatchTile *__restrict__ tiles, int n_tiles, const XtGroup *__restrict__ groups, int n_groups,
                const int32_t *__restrict__ row_q, const int32_t *__restrict__ row_k, const uint8_t *__restrict__ qexp,
                int n_tracks, unsigned long long *__restrict__ best) {
    using G = Xt<F4>;
    extern __shared__ uint8_t xsm_raw[];
    uint8_t *xsm = xsm_raw + ((128u - (smem_u32(xsm_raw) & 127u)) & 127u);
    uint8_t *r_s = xsm;                              // 2 x R_BYTES
    uint8_t *q_s = xsm + 2 * G::R_BYTES;             // STAGES x STAGE_BYTES
    constexpr int XT_STAGES = G::STAGES;
    __shared__ __align__(8) uint64_t q_full[XT_STAGES], q_empty[XT_STAGES], h0_issued[XT_STAGES], r_full[2], r_empty[2],
        acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long total = (long long)n_groups * n_tiles;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 32) {
        for (int s = 0; s < XT_STAGES; ++s) {
            mbar_init(&q_full[s], 1);
            mbar_init(&q_empty[s], 2);          // both issuing threads have read the stage
            mbar_init(&h0_issued[s], 1);        // the first half's MMAs of this stage have been issued
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&r_full[b], 128);
            mbar_init(&r_empty[b], 2);          // both halves have finished with the reference buffer
        }
        for (int h = 0; h < 2; ++h) {
            mbar_init(&acc_full[h], 1);
            mbar_init(&acc_empty[h], 128);      // the four epilogue warps of that half
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (F4) {
        // every block scale = 2^0: UE8M0 byte 0x7F in all four bytes of TMEM columns 480..511, whatever the layout
        if (warp >= 6 && warp < 10) tmem_fill_x32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + XT_SF_COL, 0x7F7F7F7Fu);
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer: the group's expanded query words, JS words per stage =====
            uint32_t sidx = 0;
            for (long long item = blockIdx.x; item < total; item += gridDim.x) {
                const XtItem it = xt_item<F4>(item, n_tiles, tiles, track_start, groups);
                if (it.need <= 0) continue;
                const uint8_t *src = qexp + groups[it.g].exp_off;
                const int nst = (it.kmax + G::JS - 1) / G::JS;
                for (int st = 0; st < nst; ++st, ++sidx) {
                    const uint32_t s = sidx % XT_STAGES;
                    if (sidx >= XT_STAGES) mbar_wait(&q_empty[s], ((sidx / XT_STAGES) - 1) & 1);
                    mbar_expect_tx(&q_full[s], G::STAGE_BYTES);
                    bulk_load_1d(q_s + s * G::STAGE_BYTES, src + (size_t)st * G::STAGE_BYTES, G::STAGE_BYTES, &q_full[s]);
                }
            }
        }
    } else if (warp == 1 || warp == 14) {
        if (lane == 0) {
            // ===== MMA issuers: warp 1 = first N half of every tile (TMEM columns [0, HALF)), warp 14 = second half =====
            // Both walk the same sequence of stages (JS query words of one K chunk of one tile). Everything a stage needs sits in
            // registers before its MMAs go out; per MMA the thread does two 64-bit adds and the instruction itself.
            const int H = warp == 1 ? 0 : 1;
            const uint64_t a_hi = xt_desc(0, G::Q_LBO, 128), b_hi = xt_desc(0, G::R_LBO, 128) + (uint64_t)(H ? G::HALF : 0);
            const uint32_t sfa = tmem_base + XT_SF_COL, sfb = tmem_base + XT_SF_COL + 16;
            const uint32_t d = tmem_base + (H ? G::HALF : 0);
            uint32_t sidx = 0, ridx = 0, cnt = 0;
            for (long long item = blockIdx.x; item < total; item += gridDim.x) {
                const XtItem it = xt_item<F4>(item, n_tiles, tiles, track_start, groups);
                if (it.need <= 0) continue;
                const int n = H ? (it.need > G::HALF ? (it.need - G::HALF + 15) & ~15 : 0) : min(G::HALF, (it.need + 15) & ~15);
                const uint32_t id = F4 ? xt_idesc_f4(n ? n : 16) : xt_idesc(n ? n : 16);
                const uint32_t sidx_last = sidx + (uint32_t)((it.kmax + G::JS - 1) / G::JS) - 1;   // last stage of this tile
                if (n) {
                    if (cnt > 0) mbar_wait(&acc_empty[H], (cnt - 1) & 1);     // the epilogue has drained this half
                    tc_fence_after();
                }
                uint32_t acc = 0;
                for (int c = 0; c < it.nchunks; ++c, ++ridx) {
                    const uint32_t rb = ridx & 1;
                    mbar_wait(&r_full[rb], (ridx >> 1) & 1);
                    tc_fence_after();
                    const int j0 = c * it.jc, j1 = min(it.kmax, j0 + it.jc);
                    const uint64_t b_base = b_hi + (uint64_t)(smem_u32(r_s + rb * G::R_BYTES) >> 4);
                    for (int j = j0; j < j1; j += G::JS, ++sidx) {
                        const uint32_t s = sidx % XT_STAGES;
                        if (H && G::LAG > 0) {
                            // stay LAG stages behind the first half (or wait for its last stage of this tile)
                            const uint32_t g = min(sidx + (uint32_t)G::LAG, sidx_last);
                            mbar_wait(&h0_issued[g % XT_STAGES], (g / XT_STAGES) & 1);
                        }
                        mbar_wait(&q_full[s], (sidx / XT_STAGES) & 1);
                        tc_fence_after();
                        if (n) {
                            const uint64_t a_base = a_hi + (uint64_t)(smem_u32(q_s + s * G::STAGE_BYTES) >> 4);
                            const uint64_t b_row = b_base + (uint64_t)(j - j0);
                            const int nj = min(G::JS, j1 - j);
                            auto one = [&](int jj) {
#pragma unroll
                                for (int kk = 0; kk < G::KSTEPS; ++kk) {
                                    const uint64_t a_desc = a_base + (uint64_t)((jj * G::QWORD_BYTES + kk * 2 * G::Q_LBO) >> 4);
                                    const uint64_t b_desc = b_row + (uint64_t)jj + (uint64_t)((kk * 2 * G::R_LBO) >> 4);
                                    if (F4) tc_mma_mxf4(d, a_desc, b_desc, id, sfa, sfb, acc);
                                    else tc_mma_i8(d, a_desc, b_desc, id, acc);
                                    acc = 1;
                                }
                            };
                            if (nj == G::JS) {
#pragma unroll
                                for (int jj = 0; jj < G::JS; ++jj) one(jj);
                            } else {
                                for (int jj = 0; jj < nj; ++jj) one(jj);
                            }
                        }
                        if (!H) mbar_arrive(&h0_issued[s]);
                        tc_commit(&q_empty[s]);        // this half is done with the stage (a commit without MMAs arrives at once)
                    }
                    tc_commit(&r_empty[rb]);
                }
                if (n) {
                    tc_commit(&acc_full[H]);
                    ++cnt;
                }
            }
        }
    } else if (warp >= 2 && warp < 6) {
        // ===== reference expanders: rows [tile_start + j0, + NOFF + (j1 - j0)) of the track, chunk-column layout =====
        const int et = threadIdx.x - 64;
        uint32_t ridx = 0;
        for (long long item = blockIdx.x; item < total; item += gridDim.x) {
            const XtItem it = xt_item<F4>(item, n_tiles, tiles, track_start, groups);
            if (it.need <= 0) continue;
            for (int c = 0; c < it.nchunks; ++c, ++ridx) {
                const uint32_t rb = ridx & 1;
                if (ridx >= 2) mbar_wait(&r_empty[rb], ((ridx >> 1) - 1) & 1);
                const int j0 = c * it.jc, j1 = min(it.kmax, j0 + it.jc);
                const int nrows = G::NOFF + (j1 - j0);
                uint4 *dst = reinterpret_cast<uint4 *>(r_s + rb * G::R_BYTES);
                const uint64_t *src = words + it.tbeg;
                const int g0 = it.tile_start + j0;
                for (int w = et; w < nrows; w += 128) {
                    const int g = g0 + w;
                    const bool valid = g < it.n_r;
                    const uint64_t x = valid ? src[g] : 0ull;
#pragma unroll
                    for (int cc = 0; cc < G::CPW; ++cc)
                        dst[cc * G::NR + w] = valid ? xt_chunk<F4>(x, cc) : make_uint4(0u, 0u, 0u, 0u);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> tcgen05 reads
                mbar_arrive(&r_full[rb]);
            }
        }
    } else if (warp >= 6 && warp < 14) {
        // ===== epilogue: 8 warps. Warp w reads TMEM lanes 32 (w % 4) .. +31 = queries; warps 6-9 own the first N half of every
        // tile (columns = offsets [0, HALF)), warps 10-13 the second half, each with its own full / empty barrier pair, so a
        // half is drained as soon as ITS last MMA has completed (the issuer runs the second half LAG stages behind the first).
        // Each thread keeps the best (dot, column) of its columns; the two partial minima of a query meet in the 64-bit
        // atomicMin, whose key carries the offset, so the lowest offset wins among equal distances. =====
        const int half = (warp - 6) >> 2;
        const int m = (warp & 3) * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t tcount = 0;
        for (long long item = blockIdx.x; item < total; item += gridDim.x) {
            const XtItem it = xt_item<F4>(item, n_tiles, tiles, track_start, groups);
            if (it.need <= 0) continue;
            const int c_base = half * G::HALF;
            if (it.need <= c_base) continue;                             // the tile has no second half (its last, short tile)
            const int q = row_q[it.g * XT_NQ + m];
            const int k_eff = min(row_k[it.g * XT_NQ + m], it.n_r);      // storage.h:34-38
            // valid columns of the tile: n <= lim; this half covers [c_base, c_base + ncol)
            const int lim = q >= 0 ? min((it.n_r - k_eff) - it.tile_start, it.need - 1) : -1;
            const int ncol = min(G::HALF, it.need - c_base);
            mbar_wait(&acc_full[half], tcount & 1);
            tc_fence_after();
            int best_d = 0, best_n = 0;      // best_d = dot + XT_BIAS of the best column so far (0: none)
#pragma unroll 1
            for (int c0 = 0; c0 < ncol; c0 += G::ECH) {
                // a load never crosses into the other half (whose accumulator is live) nor into the scale-factor columns: the
                // last chunk of a 240-column half is moved back to end at the boundary; re-evaluated columns cannot win again
                // (a column only replaces the best one with a strictly larger dot)
                const int cc = c_base + min(c0, G::HALF - G::ECH);
                const int li = lim - cc;
                if (F4) {
                    // f32 accumulators holding exact integers: key = dot * 32 + (31 - i) stays below 2^24, exact in f32
                    uint32_t v[32];
                    tmem_ld_x32(t_lane + (uint32_t)cc, v);
                    float bk = -3.0e38f;
                    if (__all_sync(0xFFFFFFFFu, li >= 31)) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) bk = fmaxf(bk, fmaf(__uint_as_float(v[i]), 32.0f, float(31 - i)));
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float key = fmaf(__uint_as_float(v[i]), 32.0f, float(31 - i));
                            bk = fmaxf(bk, i <= li ? key : -3.0e38f);
                        }
                    }
                    const int ki = __float2int_rn(fmaxf(bk, -1.0e9f));
                    const int d = bk > -1.0e38f ? (ki >> 5) + XT_BIAS : 0;
                    const bool better = d > best_d;    // strict: an equal distance at a higher offset never replaces
                    best_n = better ? cc + 31 - (ki & 31) : best_n;
                    best_d = better ? d : best_d;
                } else {
                    uint32_t v[64];
                    tmem_ld_x64(t_lane + (uint32_t)cc, v);
                    uint32_t bk = 0;
                    if (__all_sync(0xFFFFFFFFu, li >= 63)) {
#pragma unroll
                        for (int i = 0; i < 64; ++i)
                            bk = max(bk, ((uint32_t)((int)v[i] + XT_BIAS) << 6) | (uint32_t)(63 - i));
                    } else {
#pragma unroll
                        for (int i = 0; i < 64; ++i) {
                            const uint32_t key = ((uint32_t)((int)v[i] + XT_BIAS) << 6) | (uint32_t)(63 - i);
                            bk = max(bk, i <= li ? key : 0u);
                        }
                    }
                    const int d = int(bk >> 6);
                    if (d > best_d) {
                        best_d = d;
                        best_n = cc + 63 - int(bk & 63u);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[half]);
            if (best_d > 0) {
                const long long dot = (long long)best_d - XT_BIAS;
                const unsigned long long dist = (unsigned long long)((64ll * k_eff - dot) >> 1);
                atomicMin(best + (size_t)q * (size_t)n_tracks + (size_t)it.track,
                          (dist << HPFW_KEY_OFFSET_BITS) | (unsigned long long)(it.tile_start + best_n));
            }
            ++tcount;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

template <int F4>
static int match_tc_launch(hpfw_ctx *ctx, const hpfw_db *db, const uint64_t *d_qwords, const int64_t *d_qstart,
                           const XtGroup *d_groups, const int32_t *d_row_q, const int32_t *d_row_k, int n_groups,
                           int kpad_max, uint8_t *d_qexp, unsigned long long *d_best, cudaStream_t stream) {
    using G = Xt<F4>;
    const MatchTile *tiles = F4 ? db->d_tiles_f4 : db->d_tiles_tc;
    const int n_tiles = F4 ? db->n_tiles_f4 : db->n_tiles_tc;
    if (n_groups <= 0 || n_tiles <= 0) return HPFW_OK;
    {
        KernelScope ks(ctx, HPFW_K_OTHER, stream);
        xt_expand_queries_kernel<F4><<<dim3(kpad_max, n_groups), XT_NQ, 0, stream>>>(d_qwords, d_qstart, d_groups, d_row_q,
                                                                                     d_row_k, d_qexp);
        HPFW_CUDA_TRY(cudaGetLastError());
    }
    const size_t smem = 2 * (size_t)G::R_BYTES + (size_t)G::STAGES * G::STAGE_BYTES + 128;
    if (smem > size_t(ctx->max_smem_optin))
        HPFW_FAIL(HPFW_ERR_LIMIT, "match_tc: needs %zu B shared memory (> %d)", smem, ctx->max_smem_optin);
    HPFW_CUDA_TRY(cudaFuncSetAttribute(match_tc_kernel<F4>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    const long long total = (long long)n_groups * n_tiles;
    const int grid = int(std::min<long long>(total, ctx->sm_count));
    KernelScope ks(ctx, HPFW_K_MATCH_TC, stream);
    match_tc_kernel<F4><<<grid, XT_THREADS, smem, stream>>>(db->d_words, db->d_track_start, tiles, n_tiles, d_groups, n_groups,
                                                            d_row_q, d_row_k, d_qexp, db->n_tracks, d_best);
    HPFW_CUDA_TRY(cudaGetLastError());
    return HPFW_OK;
}

int match_tc_run(hpfw_ctx *ctx, const hpfw_db *db, int f4, const uint64_t *d_qwords, const int64_t *d_qstart,
                 const XtGroup *d_groups, const int32_t *d_row_q, const int32_t *d_row_k, int n_groups, int kpad_max,
                 uint8_t *d_qexp, unsigned long long *d_best, cudaStream_t stream) {
    return f4 ? match_tc_launch<1>(ctx, db, d_qwords, d_qstart, d_groups, d_row_q, d_row_k, n_groups, kpad_max, d_qexp,
                                   d_best, stream)
              : match_tc_launch<0>(ctx, db, d_qwords, d_qstart, d_groups, d_row_q, d_row_k, n_groups, kpad_max, d_qexp,
                                   d_best, stream);
}

}  // namespace hpfw_b200
