// hpfw_b200/csrc/matcher.cuh — the hashprint database object shared by the two matcher kernels (matcher.cu: XOR + POPC on
// the integer pipes; match_tc.cu: the same cross-correlation as an exact int8 GEMM on the tensor cores).
#pragma once

#include "common.cuh"

namespace hpfw_b200 {

struct MatchTile {
    int32_t track;
    int32_t start;
};

// tensor-core matcher geometry (match_tc.cu)
constexpr int XT_NQ = 128;      // queries per group = M of the MMA (TMEM lanes)
constexpr int XT_NOFF = 512;    // alignment offsets per tile = 2 x N(256) = all 512 TMEM columns
constexpr int XT_JS = 4;        // query words per TMA stage
constexpr int XT_MIN_FILL = 24; // a partial group this full is still faster on the tensor cores than on the integer pipes

// one group of up to XT_NQ queries of similar length: its expanded (s8) words start at exp_off bytes into the scratch
struct XtGroup {
    int64_t exp_off;
    int32_t kmax;   // longest query of the group (words)
    int32_t kmin;   // shortest
};

inline int xt_kpad(int kmax) { return ((kmax < 1 ? 1 : kmax) + XT_JS - 1) / XT_JS * XT_JS; }

}  // namespace hpfw_b200

struct hpfw_db {
    hpfw_ctx *ctx = nullptr;
    int n_tracks = 0;
    int64_t total_words = 0;
    int64_t track_base = 0;
    uint64_t *d_words = nullptr;
    int64_t *d_track_start = nullptr;
    hpfw_b200::MatchTile *d_tiles = nullptr;      // 2048-offset tiles (matcher.cu)
    int n_tiles = 0;
    hpfw_b200::MatchTile *d_tiles_tc = nullptr;   // 512-offset tiles (match_tc.cu)
    int n_tiles_tc = 0;
    std::vector<int64_t> offsets;  // host copy
};

namespace hpfw_b200 {
// Expands the grouped queries to s8 and runs the tensor-core matcher for n_groups groups; best[q * n_tracks + track] receives
// (dist << 20 | offset) minima exactly as match_kernel writes them. All table pointers are device pointers.
int match_tc_run(hpfw_ctx *ctx, const hpfw_db *db, const uint64_t *d_qwords, const int64_t *d_qstart,
                 const XtGroup *d_groups, const int32_t *d_row_q, const int32_t *d_row_k, int n_groups, int kpad_max,
                 uint8_t *d_qexp, unsigned long long *d_best, cudaStream_t stream);
}  // namespace hpfw_b200
