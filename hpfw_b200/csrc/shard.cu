// hpfw_b200/csrc/shard.cu — the hashprint database sharded by track over several GPUs, with NCCL called from inside the library.
//
// The reference's matcher is single-threaded (/root/reference/include/hpfw/audioproblems/live-song-id/storage.h:29 "TODO: maybe
// parallelize search"). Here (BASELINE.json north_star): the DB is partitioned into contiguous track
// ranges balanced by matcher work, the query batch is replicated, every GPU ranks its own shard into per-query top-k keys
// (dist << 40 | GLOBAL track << 20 | offset), ONE in-place ncclAllGather over NVLink exchanges the [Q][k] key arrays and
// merge_kernel keeps the k smallest of world * k per query. Unsigned '<' on keys is the reference's strict-'<' scan order, so
// the result is bit-identical for any number of shards. No other data-path collective exists on the query side; the index
// side has one ncclAllReduce(sum) of the 2420 x 2420 covariance accumulator (parallel_collector.h:94-97's mutex).
//
// Two ways to own the GPUs:
//   local: ONE process drives n devices (hpfw_shard_create_local; ncclCommInitAll) — what the C++ Storage plug-in
//          db::ShardedMemoryStorage uses, so LiveSongIdentification<Collector, ShardedMemoryStorage> spans a node from one process;
//   rank:  one process per GPU (hpfw_shard_create_rank; ncclCommInitRank with an id from rank 0's hpfw_shard_unique_id, carried to
//          the other ranks by whatever launched them) — what bench.py uses under torchrun.
// libnccl.so.2 is loaded with dlopen on first use: the library itself has no link-time dependency on NCCL, and a process that
// already carries an NCCL (PyTorch bundles one) shares that copy.
#include "matcher.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

using namespace hpfw_b200;

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
};

NcclApi g_nccl;
std::once_flag g_nccl_once;
std::string g_nccl_error;

void nccl_load() {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) {
        g_nccl_error = std::string("cannot load libnccl.so.2: ") + dlerror();
        return;
    }
    bool ok = true;
    auto sym = [&](const char *name) {
        void *p = dlsym(g_nccl.handle, name);
        if (!p) {
            ok = false;
            g_nccl_error = std::string("libnccl.so.2 lacks ") + name;
        }
        return p;
    };
    g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(sym("ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(sym("ncclCommInitRank"));
    g_nccl.CommInitAll = reinterpret_cast<decltype(g_nccl.CommInitAll)>(sym("ncclCommInitAll"));
    g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(sym("ncclCommDestroy"));
    g_nccl.AllGather = reinterpret_cast<decltype(g_nccl.AllGather)>(sym("ncclAllGather"));
    g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(sym("ncclAllReduce"));
    g_nccl.Broadcast = reinterpret_cast<decltype(g_nccl.Broadcast)>(sym("ncclBroadcast"));
    g_nccl.GroupStart = reinterpret_cast<decltype(g_nccl.GroupStart)>(sym("ncclGroupStart"));
    g_nccl.GroupEnd = reinterpret_cast<decltype(g_nccl.GroupEnd)>(sym("ncclGroupEnd"));
    g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(sym("ncclGetErrorString"));
    g_nccl.GetVersion = reinterpret_cast<decltype(g_nccl.GetVersion)>(sym("ncclGetVersion"));
    if (!ok) {
        dlclose(g_nccl.handle);
        g_nccl.handle = nullptr;
    }
}

int nccl_require() {
    std::call_once(g_nccl_once, nccl_load);
    if (!g_nccl.handle) HPFW_FAIL(HPFW_ERR_STATE, "NCCL is not available: %s", g_nccl_error.c_str());
    return HPFW_OK;
}

#define HPFW_NCCL_TRY(expr)                                                                                     \
    do {                                                                                                        \
        ncclResult_t _r = (expr);                                                                               \
        if (_r != ncclSuccess) {                                                                                \
            hpfw_b200::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r)); \
            return HPFW_ERR_CUDA;                                                                               \
        }                                                                                                       \
    } while (0)

struct ShardDev {
    hpfw_ctx *ctx = nullptr;
    bool own_ctx = false;
    ncclComm_t comm = nullptr;
    hpfw_db *db = nullptr;
    DeviceBuffer gather, q, keys, tmp;
    int64_t track_base = 0;
};

}  // namespace

struct hpfw_shard {
    bool local = false;       // one process, several devices
    int world = 1, rank = 0;  // rank mode: this process's rank; local mode: world = devs.size()
    std::vector<ShardDev> devs;
    PinnedBuffer pin_q, pin_out;
    std::vector<int> bounds;  // local mode: track range of device d = [bounds[d], bounds[d+1])
};

static void plan_bounds(const int64_t *track_words, int n_tracks, int n_shards, int query_words, std::vector<int> &bounds,
                        const double *speed = nullptr) {
    // contiguous ranges (the global track index = DB order, which the tie rule relies on) balanced by sum (n_r - k + 1) * k
    std::vector<double> csum(size_t(n_tracks) + 1, 0.0);
    for (int r = 0; r < n_tracks; ++r) {
        const int64_t n = track_words[r], k = std::min<int64_t>(n, query_words);
        csum[size_t(r) + 1] = csum[size_t(r)] + double((n - k + 1) * std::max<int64_t>(k, 1));
    }
    const double total = csum[size_t(n_tracks)];
    // shard s gets the share speed[s] / sum(speed) of the work (equal shares without speeds): GPUs of one node do not run at
    // the same clock under their power caps, and a collective waits for the slowest rank
    double speed_sum = 0.0, speed_acc = 0.0;
    for (int s = 0; s < n_shards; ++s) speed_sum += speed ? std::max(speed[s], 1e-9) : 1.0;
    bounds.assign(1, 0);
    for (int s = 1; s < n_shards; ++s) {
        speed_acc += speed ? std::max(speed[s - 1], 1e-9) : 1.0;
        const double target = total * speed_acc / speed_sum;
        int b = int(std::lower_bound(csum.begin(), csum.end(), target) - csum.begin());
        if (b > 0 && std::fabs(csum[size_t(b) - 1] - target) <= std::fabs(csum[size_t(std::min(b, n_tracks))] - target)) --b;
        bounds.push_back(std::min(std::max(b, bounds.back()), n_tracks));
    }
    bounds.push_back(n_tracks);
}

extern "C" {

int hpfw_shard_plan(const int64_t *track_words, int n_tracks, int n_shards, int query_words, int *bounds_out) {
    if (n_tracks < 0 || n_shards < 1 || !bounds_out || (n_tracks > 0 && !track_words))
        HPFW_FAIL(HPFW_ERR_ARG, "hpfw_shard_plan: bad argument");
    std::vector<int> b;
    plan_bounds(track_words, n_tracks, n_shards, query_words > 0 ? query_words : 385, b);
    memcpy(bounds_out, b.data(), sizeof(int) * b.size());
    return HPFW_OK;
}

int hpfw_shard_plan_weighted(const int64_t *track_words, int n_tracks, int n_shards, int query_words, const double *speed,
                             int *bounds_out) {
    if (n_tracks < 0 || n_shards < 1 || !bounds_out || !speed || (n_tracks > 0 && !track_words))
        HPFW_FAIL(HPFW_ERR_ARG, "hpfw_shard_plan_weighted: bad argument");
    std::vector<int> b;
    plan_bounds(track_words, n_tracks, n_shards, query_words > 0 ? query_words : 385, b, speed);
    memcpy(bounds_out, b.data(), sizeof(int) * b.size());
    return HPFW_OK;
}

int hpfw_shard_unique_id(void *id128_out) {
    if (!id128_out) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_shard_unique_id: NULL argument");
    HPFW_TRY(nccl_require());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    HPFW_NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(id128_out, &id, sizeof(id));
    return HPFW_OK;
}

int hpfw_shard_nccl_version(void) {
    if (nccl_require() != HPFW_OK) return 0;
    int v = 0;
    return g_nccl.GetVersion(&v) == ncclSuccess ? v : 0;
}

int hpfw_shard_create_rank(hpfw_ctx *ctx, int rank, int world, const void *id128, hpfw_shard **out) {
    if (!ctx || !out || world < 1 || rank < 0 || rank >= world || (world > 1 && !id128))
        HPFW_FAIL(HPFW_ERR_ARG, "hpfw_shard_create_rank: bad argument");
    *out = nullptr;
    DeviceGuard g(ctx->device);
    hpfw_shard *s = new hpfw_shard();
    s->local = false;
    s->world = world;
    s->rank = rank;
    s->devs.resize(1);
    s->devs[0].ctx = ctx;
    if (world > 1) {
        int st = nccl_require();
        if (st == HPFW_OK) {
            ncclUniqueId id;
            memcpy(&id, id128, sizeof(id));
            ncclResult_t r = g_nccl.CommInitRank(&s->devs[0].comm, world, id, rank);
            if (r != ncclSuccess) {
                set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
                st = HPFW_ERR_CUDA;
            }
        }
        if (st != HPFW_OK) {
            delete s;
            return st;
        }
    }
    *out = s;
    return HPFW_OK;
}

int hpfw_shard_create_local(const int *devices, int n_devices, hpfw_shard **out) {
    if (!out || n_devices < 1 || n_devices > 64) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_shard_create_local: bad argument");
    *out = nullptr;
    hpfw_shard *s = new hpfw_shard();
    s->local = true;
    s->world = n_devices;
    s->devs.resize(size_t(n_devices));
    std::vector<int> devlist(static_cast<size_t>(n_devices));
    int st = HPFW_OK;
    for (int d = 0; d < n_devices && st == HPFW_OK; ++d) {
        devlist[size_t(d)] = devices ? devices[d] : d;
        st = hpfw_ctx_create(devlist[size_t(d)], &s->devs[size_t(d)].ctx);
        s->devs[size_t(d)].own_ctx = st == HPFW_OK;
    }
    if (st == HPFW_OK && n_devices > 1) {
        st = nccl_require();
        if (st == HPFW_OK) {
            std::vector<ncclComm_t> comms(static_cast<size_t>(n_devices));
            ncclResult_t r = g_nccl.CommInitAll(comms.data(), n_devices, devlist.data());
            if (r != ncclSuccess) {
                set_error("ncclCommInitAll failed: %s", g_nccl.GetErrorString(r));
                st = HPFW_ERR_CUDA;
            } else {
                for (int d = 0; d < n_devices; ++d) s->devs[size_t(d)].comm = comms[size_t(d)];
            }
        }
    }
    if (st != HPFW_OK) {
        hpfw_shard_destroy(s);
        return st;
    }
    *out = s;
    return HPFW_OK;
}

void hpfw_shard_destroy(hpfw_shard *s) {
    if (!s) return;
    for (auto &d : s->devs) {
        if (!d.ctx) continue;
        DeviceGuard g(d.ctx->device);
        cudaDeviceSynchronize();
        if (d.db) hpfw_db_destroy(d.db);
        if (d.comm) g_nccl.CommDestroy(d.comm);
        d.gather.release();
        d.q.release();
        d.keys.release();
        d.tmp.release();
        if (d.own_ctx) hpfw_ctx_destroy(d.ctx);
    }
    s->pin_q.release();
    s->pin_out.release();
    delete s;
}

int hpfw_shard_world(const hpfw_shard *s) { return s ? s->world : 0; }
int hpfw_shard_rank(const hpfw_shard *s) { return s ? (s->local ? 0 : s->rank) : -1; }
hpfw_ctx *hpfw_shard_ctx(hpfw_shard *s, int local_index) {
    if (!s || local_index < 0 || local_index >= int(s->devs.size())) return nullptr;
    return s->devs[size_t(local_index)].ctx;
}
hpfw_db *hpfw_shard_db(hpfw_shard *s, int local_index) {
    if (!s || local_index < 0 || local_index >= int(s->devs.size())) return nullptr;
    return s->devs[size_t(local_index)].db;
}

// ------------------------------------------------------------------------------------------------------------- build
int hpfw_shard_build_rank(hpfw_shard *s, const uint64_t *words, const int64_t *offsets, int n_tracks, int64_t track_base) {
    if (!s || s->local) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_shard_build_rank: needs a rank-mode shard");
    ShardDev &d = s->devs[0];
    if (d.db) hpfw_db_destroy(d.db);
    d.db = nullptr;
    d.track_base = track_base;
    return hpfw_db_build(d.ctx, words, offsets, n_tracks, track_base, &d.db);
}

int hpfw_shard_build_rank_device(hpfw_shard *s, const uint64_t *d_words, const int64_t *offsets, int n_tracks,
                                 int64_t track_base, void *stream) {
    if (!s || s->local) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_shard_build_rank_device: needs a rank-mode shard");
    ShardDev &d = s->devs[0];
    if (d.db) hpfw_db_destroy(d.db);
    d.db = nullptr;
    d.track_base = track_base;
    return hpfw_db_build_device(d.ctx, d_words, offsets, n_tracks, track_base, stream, &d.db);
}

int hpfw_shard_build(hpfw_shard *s, const uint64_t *words, const int64_t *offsets, int n_tracks, int query_words_hint) {
    if (!s || !s->local || n_tracks < 0 || (n_tracks > 0 && !offsets)) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_shard_build: bad argument");
    std::vector<int64_t> lens(static_cast<size_t>(n_tracks));
    for (int r = 0; r < n_tracks; ++r) lens[size_t(r)] = offsets[r + 1] - offsets[r];
    plan_bounds(lens.data(), n_tracks, s->world, query_words_hint > 0 ? query_words_hint : 385, s->bounds);
    static const int64_t zero[1] = {0};
    for (int dv = 0; dv < s->world; ++dv) {
        ShardDev &d = s->devs[size_t(dv)];
        if (d.db) hpfw_db_destroy(d.db);
        d.db = nullptr;
        const int lo = s->bounds[size_t(dv)], hi = s->bounds[size_t(dv) + 1];
        d.track_base = lo;
        HPFW_TRY(hpfw_db_build(d.ctx, words, n_tracks ? offsets + lo : zero, hi - lo, lo, &d.db));
    }
    return HPFW_OK;
}

int hpfw_shard_build_device(hpfw_shard *s, const uint64_t *d_words, const int64_t *src_offsets, const int64_t *lengths,
                            int n_tracks, int query_words_hint) {
    if (!s || !s->local || n_tracks < 0 || (n_tracks > 0 && (!src_offsets || !lengths)))
        HPFW_FAIL(HPFW_ERR_ARG, "hpfw_shard_build_device: bad argument");
    plan_bounds(lengths, n_tracks, s->world, query_words_hint > 0 ? query_words_hint : 385, s->bounds);
    ShardDev &d0 = s->devs[0];
    for (int dv = 0; dv < s->world; ++dv) {
        ShardDev &d = s->devs[size_t(dv)];
        if (d.db) hpfw_db_destroy(d.db);
        d.db = nullptr;
        const int lo = s->bounds[size_t(dv)], hi = s->bounds[size_t(dv) + 1];
        d.track_base = lo;
        if (dv == 0) {
            HPFW_TRY(hpfw_db_build_gather_device(d.ctx, d_words, src_offsets + lo, lengths + lo, hi - lo, lo, nullptr, &d.db));
            continue;
        }
        // device 0 gathers the range into DB order, the range crosses NVLink once, the peer builds its shard from it
        hpfw_db *stage = nullptr;
        HPFW_TRY(hpfw_db_build_gather_device(d0.ctx, d_words, src_offsets + lo, lengths + lo, hi - lo, 0, nullptr, &stage));
        const int64_t nw = stage->total_words;
        int st = HPFW_OK;
        {
            DeviceGuard g(d.ctx->device);
            st = d.tmp.reserve(sizeof(uint64_t) * size_t(std::max<int64_t>(nw, 1)));
            if (st == HPFW_OK && nw > 0 &&
                cudaMemcpyPeer(d.tmp.ptr, d.ctx->device, stage->d_words, d0.ctx->device, sizeof(uint64_t) * size_t(nw)) != cudaSuccess) {
                set_error("hpfw_shard_build_device: peer copy to device %d failed: %s", d.ctx->device,
                          cudaGetErrorString(cudaGetLastError()));
                st = HPFW_ERR_CUDA;
            }
            if (st == HPFW_OK)
                st = hpfw_db_build_device(d.ctx, d.tmp.as<uint64_t>(), stage->offsets.data(), hi - lo, lo, nullptr, &d.db);
            d.tmp.release();
        }
        hpfw_db_destroy(stage);
        HPFW_TRY(st);
    }
    return HPFW_OK;
}

// ------------------------------------------------------------------------------------------------------------- match
static int shard_gather_merge(hpfw_shard *s, ShardDev &d, int nq, int topk, uint64_t *d_keys_out, cudaStream_t stream) {
    const size_t cnt = size_t(nq) * size_t(topk);
    if (s->world > 1)
        HPFW_NCCL_TRY(g_nccl.AllGather(d.gather.as<uint64_t>() + size_t(s->local ? (&d - s->devs.data()) : s->rank) * cnt,
                                       d.gather.ptr, cnt, ncclUint64, d.comm, stream));
    if (d_keys_out)
        HPFW_TRY(hpfw_topk_merge_device(d.ctx, d.gather.as<uint64_t>(), s->world, nq, topk, d_keys_out, stream));
    return HPFW_OK;
}

int hpfw_shard_match_device(hpfw_shard *s, const uint64_t *d_qwords, const int64_t *qoffsets, int n_queries, int topk,
                            uint64_t *d_keys_out, void *stream_) {
    if (!s || s->local || !qoffsets || !d_keys_out || n_queries < 0 || topk < 1)
        HPFW_FAIL(HPFW_ERR_ARG, "hpfw_shard_match_device: bad argument (needs a rank-mode shard)");
    if (n_queries == 0) return HPFW_OK;
    ShardDev &d = s->devs[0];
    if (!d.db) HPFW_FAIL(HPFW_ERR_STATE, "hpfw_shard_match_device: no shard built on this rank");
    DeviceGuard g(d.ctx->device);
    cudaStream_t stream = d.ctx->pick(stream_);
    const size_t cnt = size_t(n_queries) * size_t(topk);
    HPFW_TRY(d.gather.reserve(sizeof(uint64_t) * cnt * size_t(s->world)));
    // the local top-k lands in this rank's slot of the gather buffer: the all-gather is in place
    HPFW_TRY(hpfw_db_match_device(d.db, d_qwords, qoffsets, n_queries, topk, d.gather.as<uint64_t>() + size_t(s->rank) * cnt, stream));
    return shard_gather_merge(s, d, n_queries, topk, d_keys_out, stream);
}

// local mode: the queries are on every device's q buffer already
static int shard_match_local(hpfw_shard *s, const int64_t *qoffsets, int nq, int topk, hpfw_match *out) {
    const size_t cnt = size_t(nq) * size_t(topk);
    for (auto &d : s->devs) {
        if (!d.db) HPFW_FAIL(HPFW_ERR_STATE, "hpfw_shard: build has not been called");
        DeviceGuard g(d.ctx->device);
        HPFW_TRY(d.gather.reserve(sizeof(uint64_t) * cnt * size_t(s->world)));
        const size_t me = size_t(&d - s->devs.data());
        HPFW_TRY(hpfw_db_match_device(d.db, d.q.as<uint64_t>(), qoffsets, nq, topk, d.gather.as<uint64_t>() + me * cnt,
                                      d.ctx->stream));
    }
    ShardDev &d0 = s->devs[0];
    {
        DeviceGuard g0(d0.ctx->device);
        HPFW_TRY(d0.keys.reserve(sizeof(uint64_t) * cnt));
        HPFW_TRY(s->pin_out.reserve(sizeof(uint64_t) * cnt));
    }
    // all-gathers of all devices as one NCCL group (a single thread drives every communicator); NCCL enqueues them at
    // GroupEnd, so the merge on device 0 is enqueued after that to follow its all-gather in stream order
    if (s->world > 1) {
        HPFW_NCCL_TRY(g_nccl.GroupStart());
        int st = HPFW_OK;
        for (auto &d : s->devs) {
            DeviceGuard g(d.ctx->device);
            const int r = shard_gather_merge(s, d, nq, topk, nullptr, d.ctx->stream);
            if (r != HPFW_OK) st = r;
        }
        HPFW_NCCL_TRY(g_nccl.GroupEnd());
        HPFW_TRY(st);
    }
    {
        DeviceGuard g0(d0.ctx->device);
        HPFW_TRY(hpfw_topk_merge_device(d0.ctx, d0.gather.as<uint64_t>(), s->world, nq, topk, d0.keys.as<uint64_t>(), d0.ctx->stream));
    }
    DeviceGuard g0(d0.ctx->device);
    HPFW_CUDA_TRY(cudaMemcpyAsync(s->pin_out.ptr, d0.keys.ptr, sizeof(uint64_t) * cnt, cudaMemcpyDeviceToHost, d0.ctx->stream));
    HPFW_CUDA_TRY(cudaStreamSynchronize(d0.ctx->stream));
    hpfw_keys_decode(s->pin_out.as<uint64_t>(), int(cnt), out);
    return HPFW_OK;
}

int hpfw_shard_find_topk(hpfw_shard *s, const uint64_t *qwords, const int64_t *qoffsets, int n_queries, int topk,
                         hpfw_match *out) {
    if (!s || !s->local || !qoffsets || !out || n_queries < 0 || topk < 1) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_shard_find_topk: bad argument");
    if (n_queries == 0) return HPFW_OK;
    const size_t nw = size_t(qoffsets[n_queries] - qoffsets[0]);
    if (nw && !qwords) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_shard_find_topk: qwords is NULL");
    std::vector<int64_t> rel(size_t(n_queries) + 1);
    for (int q = 0; q <= n_queries; ++q) rel[size_t(q)] = qoffsets[q] - qoffsets[0];
    // one staging copy to pinned memory, then one H2D per device (the batch is small: 394 KB per 128 six-second queries)
    HPFW_TRY(s->pin_q.reserve(sizeof(uint64_t) * std::max<size_t>(nw, 1)));
    for (auto &d : s->devs) {      // a previous call's uploads must have drained before the staging buffer is rewritten
        DeviceGuard g(d.ctx->device);
        HPFW_CUDA_TRY(cudaStreamSynchronize(d.ctx->stream));
    }
    if (nw) memcpy(s->pin_q.ptr, qwords + qoffsets[0], sizeof(uint64_t) * nw);
    for (auto &d : s->devs) {
        DeviceGuard g(d.ctx->device);
        d.ctx->order_on(d.ctx->stream);
        HPFW_TRY(d.q.reserve(sizeof(uint64_t) * std::max<size_t>(nw, 1)));
        if (nw) HPFW_CUDA_TRY(cudaMemcpyAsync(d.q.ptr, s->pin_q.ptr, sizeof(uint64_t) * nw, cudaMemcpyHostToDevice, d.ctx->stream));
    }
    return shard_match_local(s, rel.data(), n_queries, topk, out);
}

int hpfw_shard_find_topk_device(hpfw_shard *s, const uint64_t *d_qwords_dev0, const int64_t *qoffsets, int n_queries, int topk,
                                hpfw_match *out) {
    if (!s || !s->local || !qoffsets || !out || n_queries < 0 || topk < 1)
        HPFW_FAIL(HPFW_ERR_ARG, "hpfw_shard_find_topk_device: bad argument");
    if (n_queries == 0) return HPFW_OK;
    const size_t nw = size_t(qoffsets[n_queries] - qoffsets[0]);
    if (nw && !d_qwords_dev0) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_shard_find_topk_device: d_qwords is NULL");
    std::vector<int64_t> rel(size_t(n_queries) + 1);
    for (int q = 0; q <= n_queries; ++q) rel[size_t(q)] = qoffsets[q] - qoffsets[0];
    // the hashprints were extracted on device 0: one broadcast over NVLink hands them to the other devices
    for (auto &d : s->devs) {
        DeviceGuard g(d.ctx->device);
        d.ctx->order_on(d.ctx->stream);
        HPFW_TRY(d.q.reserve(sizeof(uint64_t) * std::max<size_t>(nw, 1)));
    }
    ShardDev &d0 = s->devs[0];
    if (nw) {
        DeviceGuard g0(d0.ctx->device);
        HPFW_CUDA_TRY(cudaMemcpyAsync(d0.q.ptr, d_qwords_dev0 + qoffsets[0], sizeof(uint64_t) * nw, cudaMemcpyDeviceToDevice,
                                      d0.ctx->stream));
        if (s->world > 1) {
            HPFW_NCCL_TRY(g_nccl.GroupStart());
            ncclResult_t r = ncclSuccess;
            for (auto &d : s->devs) {
                DeviceGuard g(d.ctx->device);
                const ncclResult_t rr = g_nccl.Broadcast(d.q.ptr, d.q.ptr, nw, ncclUint64, 0, d.comm, d.ctx->stream);
                if (rr != ncclSuccess) r = rr;
            }
            HPFW_NCCL_TRY(g_nccl.GroupEnd());
            HPFW_NCCL_TRY(r);
        }
    }
    return shard_match_local(s, rel.data(), n_queries, topk, out);
}

// ------------------------------------------------------------------------------------- collectives of the index side
int hpfw_shard_allreduce_cov(hpfw_shard *s, void *stream_) {
    if (!s || s->local) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_shard_allreduce_cov: needs a rank-mode shard");
    ShardDev &d = s->devs[0];
    DeviceGuard g(d.ctx->device);
    cudaStream_t stream = d.ctx->pick(stream_);
    if (!d.ctx->cov_accum.ptr) HPFW_TRY(hpfw_cov_reset(d.ctx));
    if (s->world == 1) return HPFW_OK;
    // in place on the accumulator: parallel_collector.h:94-97's `accum_cov += cov` under a mutex, across GPUs
    HPFW_NCCL_TRY(g_nccl.AllReduce(d.ctx->cov_accum.ptr, d.ctx->cov_accum.ptr, size_t(HPFW_FRAME_SIZE) * HPFW_FRAME_SIZE, ncclFloat,
                                   ncclSum, d.comm, stream));
    return HPFW_OK;
}

int hpfw_shard_allgather_device(hpfw_shard *s, const void *d_send, void *d_recv, size_t bytes_per_rank, void *stream_) {
    if (!s || s->local || !d_send || !d_recv) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_shard_allgather_device: bad argument");
    ShardDev &d = s->devs[0];
    DeviceGuard g(d.ctx->device);
    cudaStream_t stream = d.ctx->pick(stream_);
    if (s->world == 1) {
        if (d_send != static_cast<char *>(d_recv))
            HPFW_CUDA_TRY(cudaMemcpyAsync(d_recv, d_send, bytes_per_rank, cudaMemcpyDeviceToDevice, stream));
        return HPFW_OK;
    }
    HPFW_NCCL_TRY(g_nccl.AllGather(d_send, d_recv, bytes_per_rank, ncclUint8, d.comm, stream));
    return HPFW_OK;
}

int hpfw_shard_allgatherv_device(hpfw_shard *s, const void *d_send, void *d_recv, const size_t *bytes_per_rank, void *stream_) {
    if (!s || s->local || !d_recv || !bytes_per_rank) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_shard_allgatherv_device: bad argument");
    ShardDev &d = s->devs[0];
    DeviceGuard g(d.ctx->device);
    cudaStream_t stream = d.ctx->pick(stream_);
    // ranks contribute different byte counts (their share of the query hashprints): one broadcast per rank into its slice of
    // the contiguous result, issued as ONE NCCL group
    size_t displ = 0, mine = 0;
    for (int r = 0; r < s->rank; ++r) mine += bytes_per_rank[r];
    if (bytes_per_rank[s->rank] && d_send && d_send != static_cast<char *>(d_recv) + mine)
        HPFW_CUDA_TRY(cudaMemcpyAsync(static_cast<char *>(d_recv) + mine, d_send, bytes_per_rank[s->rank], cudaMemcpyDeviceToDevice,
                                      stream));
    if (s->world == 1) return HPFW_OK;
    HPFW_NCCL_TRY(g_nccl.GroupStart());
    ncclResult_t res = ncclSuccess;
    for (int r = 0; r < s->world; ++r) {
        if (bytes_per_rank[r]) {
            char *seg = static_cast<char *>(d_recv) + displ;
            const ncclResult_t rr = g_nccl.Broadcast(seg, seg, bytes_per_rank[r], ncclUint8, r, d.comm, stream);
            if (rr != ncclSuccess) res = rr;
        }
        displ += bytes_per_rank[r];
    }
    HPFW_NCCL_TRY(g_nccl.GroupEnd());
    HPFW_NCCL_TRY(res);
    return HPFW_OK;
}

int hpfw_shard_broadcast_device(hpfw_shard *s, void *d_buf, size_t bytes, int root, void *stream_) {
    if (!s || s->local || !d_buf || root < 0 || root >= s->world) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_shard_broadcast_device: bad argument");
    if (s->world == 1) return HPFW_OK;
    ShardDev &d = s->devs[0];
    DeviceGuard g(d.ctx->device);
    HPFW_NCCL_TRY(g_nccl.Broadcast(d_buf, d_buf, bytes, ncclUint8, root, d.comm, d.ctx->pick(stream_)));
    return HPFW_OK;
}

}  // extern "C"
