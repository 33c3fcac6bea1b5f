// hpfw_b200/csrc/matcher.cuh — the hashprint database object shared by the two matcher kernels (matcher.cu: XOR + POPC on
// the integer pipes; match_tc.cu: the same cross-correlation as an exact +-1 GEMM on the tensor cores, fp4 or int8 operands).
#pragma once

#include "common.cuh"

namespace hpfw_b200 {

struct MatchTile {
    int32_t track;
    int32_t start;
};

// tensor-core matcher geometry (match_tc.cu)
constexpr int XT_NQ = 128;      // queries per group = M of the MMA (TMEM lanes)
constexpr int XT_NOFF = 512;    // int8 kernel: alignment offsets per tile = 2 x N(256) = all 512 TMEM columns
constexpr int XT_NOFF_F4 = 480; // fp4 kernel: 2 x N(240); 32 TMEM columns are left for the block scale factors
constexpr int XT_JS = 4;        // int8 kernel: query words per TMA stage (64 bytes per word and query)
#ifndef XT_JS_F4_N
#define XT_JS_F4_N 8
#endif
constexpr int XT_JS_F4 = XT_JS_F4_N;     // fp4 kernel (32 bytes per word and query): 32 KB stages

// one group of up to XT_NQ queries of similar length: its expanded words start at exp_off bytes into the scratch
struct XtGroup {
    int64_t exp_off;
    int32_t kmax;   // longest query of the group (words)
    int32_t kmin;   // shortest
};

inline int xt_js(int f4) { return f4 ? XT_JS_F4 : XT_JS; }
inline int xt_kpad(int kmax, int f4) { return ((kmax < 1 ? 1 : kmax) + xt_js(f4) - 1) / xt_js(f4) * xt_js(f4); }
inline int xt_word_bytes(int f4) { return f4 ? 32 : 64; }   // expanded bytes per query word

}  // namespace hpfw_b200

struct hpfw_db {
    hpfw_ctx *ctx = nullptr;
    int device = 0;          // = ctx->device, kept here so that destroying a DB after its context is not a wild read
    int n_tracks = 0;
    int64_t total_words = 0;
    int64_t track_base = 0;
    uint64_t *d_words = nullptr;
    int64_t *d_track_start = nullptr;
    hpfw_b200::MatchTile *d_tiles = nullptr;      // 2048-offset tiles (matcher.cu)
    int n_tiles = 0;
    hpfw_b200::MatchTile *d_tiles_tc = nullptr;   // 512-offset tiles (match_tc.cu, int8)
    int n_tiles_tc = 0;
    hpfw_b200::MatchTile *d_tiles_f4 = nullptr;   // 480-offset tiles (match_tc.cu, fp4)
    int n_tiles_f4 = 0;
    std::vector<int64_t> offsets;  // host copy
};

namespace hpfw_b200 {
// Expands the grouped queries to s8 (f4 = 0) or e2m1 (f4 = 1) and runs the tensor-core matcher for n_groups groups;
// best[q * n_tracks + track] receives (dist << 20 | offset) minima exactly as match_kernel writes them. All table pointers
// are device pointers. route_queries (matcher.cu) decides which queries come here.
int match_tc_run(hpfw_ctx *ctx, const hpfw_db *db, int f4, const uint64_t *d_qwords, const int64_t *d_qstart,
                 const XtGroup *d_groups, const int32_t *d_row_q, const int32_t *d_row_k, int n_groups, int kpad_max,
                 uint8_t *d_qexp, unsigned long long *d_best, cudaStream_t stream);
}  // namespace hpfw_b200
