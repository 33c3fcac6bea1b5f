// include/hpfw/audioproblems/live-song-id/sharded_storage.h — db::ShardedMemoryStorage: the MemoryStorage interface
// (/root/reference/include/hpfw/audioproblems/live-song-id/storage.h:8-93) over ALL GPUs of a node, from one process.
//
//   hpfw::LiveSongIdentification<hpfw::DefaultLiveIdCollector, hpfw::db::ShardedMemoryStorage<hpfw::DefaultLiveIdCollector>> liveid;
//
// is the whole change a user of the reference makes to spread the database over 8 B200s: the storage is a template plug-in
// point of LiveSongIdentification (live_song_id.h:18). The database is partitioned by track into contiguous ranges balanced by
// matcher work, the query batch is replicated (broadcast over NVLink when it was extracted on device 0), every GPU ranks its
// shard, one in-place ncclAllGather of the per-query top-k keys and a merge kernel produce the result — all inside the
// library (hpfw_shard_*, hpfw_b200/csrc/shard.cu). Results are bit-identical to db::MemoryStorage for any number of GPUs:
// keys (dist, global track, offset) order like the reference's strict '<' scan (storage.h:50-60).
// Devices: all visible ones, or HPFW_DEVICES="0,1,2,3" / the constructor's list. Device 0 of the list must be the collector's.
#pragma once

#include <cstdint>
#include <cstdlib>
#include <functional>
#include <limits>
#include <optional>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "../../device.h"
#include "../../io/cereal_compat.h"

namespace hpfw::db {

template <typename Collector>
class ShardedMemoryStorage {
public:
    using Hashprint = typename Collector::Hashprint;
    using Pair = typename Collector::FilenameFingerprintPair;

    struct SearchResult {        // storage.h:11-15
        std::string filename;
        size_t cnt;
        int64_t offset;
    };

    ShardedMemoryStorage() : ShardedMemoryStorage(default_devices()) {}
    explicit ShardedMemoryStorage(const std::vector<int> &devices) {
        device::check(hpfw_shard_create_local(devices.empty() ? nullptr : devices.data(),
                                              devices.empty() ? 1 : static_cast<int>(devices.size()), &shard));
    }
    ~ShardedMemoryStorage() { hpfw_shard_destroy(shard); }
    ShardedMemoryStorage(const ShardedMemoryStorage &) = delete;
    ShardedMemoryStorage &operator=(const ShardedMemoryStorage &) = delete;

    int gpus() const { return hpfw_shard_world(shard); }

    /// DB order = order of `fingerprints` (storage.h:21-25); split over the GPUs by hpfw_shard_plan.
    template <typename PairVector>
    void build(PairVector &&fingerprints) {
        names.clear();
        std::vector<uint64_t> words;
        std::vector<int64_t> offs{0};
        for (auto &p : fingerprints) {
            names.push_back(p.filename);
            words.insert(words.end(), p.fingerprint.begin(), p.fingerprint.end());
            offs.push_back(static_cast<int64_t>(words.size()));
        }
        device::check(hpfw_shard_build(shard, words.data(), offs.data(), static_cast<int>(offs.size()) - 1, 0));
        pending_words = nullptr;
        host_words = std::move(words);
        host_offs = std::move(offs);
        built = true;
    }

    /// build() from hashprints still resident on the collector's GPU (Collector::prepare_device): device 0 keeps its range, the
    /// other ranges cross NVLink once; nothing visits the host.
    template <typename DeviceHashprints>
    void build_device(const DeviceHashprints &d) {
        names = d.names;
        host_words.clear();
        host_offs.assign(1, 0);
        for (int w : d.words) host_offs.push_back(host_offs.back() + w);
        const int n = static_cast<int>(d.order.size());
        std::vector<int64_t> all_off(static_cast<size_t>(hpfw_xs_tracks(d.xs->get()))), all_len(all_off.size());
        const uint64_t *d_words = nullptr;
        {
            std::scoped_lock l(d.ctx->mutex());
            device::check(hpfw_xs_hashprints_device(d.xs->get(), &d_words, all_off.data(), all_len.data()));
        }
        std::vector<int64_t> src(static_cast<size_t>(n)), len(static_cast<size_t>(n));
        for (int i = 0; i < n; ++i) {
            src[static_cast<size_t>(i)] = all_off[static_cast<size_t>(d.order[static_cast<size_t>(i)])];
            len[static_cast<size_t>(i)] = all_len[static_cast<size_t>(d.order[static_cast<size_t>(i)])];
            if (src[static_cast<size_t>(i)] < 0) throw Error(HPFW_ERR_STATE, "ShardedMemoryStorage::build_device: unhashed track");
        }
        {
            std::scoped_lock l(d.ctx->mutex());
            device::check(hpfw_ctx_synchronize(d.ctx->get()));     // the hashing launches of the collector's context are complete
            device::check(hpfw_shard_build_device(shard, d_words, src.data(), len.data(), n, 0));
        }
        pending_words = [xs = d.xs, gen = d.generation, order = d.order, words = d.words, ctx = d.ctx]() {
            if (xs->generation != gen)
                throw Error(HPFW_ERR_STATE, "ShardedMemoryStorage::save: the collector has started another batch since "
                                            "build_device(); save() right after index()");
            std::vector<uint64_t> out;
            size_t total = 0, pos = 0;
            for (int w : words) total += static_cast<size_t>(w);
            out.resize(total);
            std::scoped_lock l2(ctx->mutex());
            for (size_t i = 0; i < order.size(); ++i) {
                device::check(hpfw_xs_hashprint_host(xs->get(), order[i], out.data() + pos));
                pos += static_cast<size_t>(words[i]);
            }
            return out;
        };
        built = true;
    }

    /// MemoryStorage::find (storage.h:27-64).
    SearchResult find(const Hashprint &hp) const {
        const int64_t qo[2] = {0, static_cast<int64_t>(hp.size())};
        hpfw_match m;
        device::check(hpfw_shard_find_topk(require(), hp.data(), qo, 1, 1, &m));
        return to_result(m);
    }

    /// Batched top-k: out[q][r] = r-th best track of query q (distance, then DB index).
    std::vector<std::vector<SearchResult>> find_topk(const std::vector<Hashprint> &queries, int topk) const {
        std::vector<uint64_t> qw;
        std::vector<int64_t> qo{0};
        for (const auto &q : queries) {
            qw.insert(qw.end(), q.begin(), q.end());
            qo.push_back(static_cast<int64_t>(qw.size()));
        }
        std::vector<hpfw_match> m(queries.size() * static_cast<size_t>(topk));
        device::check(hpfw_shard_find_topk(require(), qw.data(), qo.data(), static_cast<int>(queries.size()), topk, m.data()));
        return to_results(m, queries.size(), topk);
    }

    /// Queries whose hashprints are still in HBM on the collector's GPU (Collector::calc_hashprints_device).
    template <typename DeviceHashprints>
    std::vector<std::vector<SearchResult>> find_topk_device(const DeviceHashprints &d, int topk) const {
        const size_t nq = d.order.size();
        if (nq == 0) return {};
        const int nt = hpfw_xs_tracks(d.xs->get());
        std::vector<int64_t> off(static_cast<size_t>(nt)), len(static_cast<size_t>(nt));
        const uint64_t *d_words = nullptr;
        std::vector<int64_t> qo(nq + 1, 0);
        std::vector<hpfw_match> m(nq * static_cast<size_t>(topk));
        {
            std::scoped_lock l(d.ctx->mutex());
            device::check(hpfw_xs_hashprints_device(d.xs->get(), &d_words, off.data(), len.data()));
            for (size_t q = 0; q < nq; ++q) {
                // calc_hashprints_device hashes every track of the stream, in stream order: the store is the query batch
                if (off[static_cast<size_t>(d.order[q])] != qo[q])
                    throw Error(HPFW_ERR_STATE, "ShardedMemoryStorage::find_topk_device: query hashprints are not contiguous");
                qo[q + 1] = qo[q] + len[static_cast<size_t>(d.order[q])];
            }
            device::check(hpfw_ctx_synchronize(d.ctx->get()));
            device::check(hpfw_shard_find_topk_device(require(), d_words, qo.data(), static_cast<int>(nq), topk, m.data()));
        }
        return to_results(m, nq, topk);
    }

    /// cereal-compatible dump of the DB (storage.h:67-76).
    std::string save(const std::optional<std::string> &filename) const {
        const std::string dump_name = filename.value_or("db/dump.cereal");
        if (pending_words) {
            host_words = pending_words();
            pending_words = nullptr;
        }
        std::vector<io::NamedHashprint> v;
        for (size_t r = 0; r < names.size(); ++r)
            v.emplace_back(names[r], std::vector<uint64_t>(host_words.begin() + host_offs[r], host_words.begin() + host_offs[r + 1]));
        io::save_db(dump_name, v);
        return dump_name;
    }

    ShardedMemoryStorage &load(const std::string &dump_name) {   // storage.h:79-86
        std::vector<Pair> pairs;
        for (auto &e : io::load_db(dump_name)) pairs.push_back({std::move(e.first), std::move(e.second)});
        build(std::move(pairs));
        return *this;
    }

    size_t size() const { return names.size(); }

private:
    hpfw_shard *shard = nullptr;
    bool built = false;
    std::vector<std::string> names;
    mutable std::vector<uint64_t> host_words;
    std::vector<int64_t> host_offs;
    mutable std::function<std::vector<uint64_t>()> pending_words;

    static std::vector<int> default_devices() {
        std::vector<int> devs;
        if (const char *env = std::getenv("HPFW_DEVICES")) {
            std::stringstream ss(env);
            std::string tok;
            while (std::getline(ss, tok, ','))
                if (!tok.empty()) devs.push_back(std::atoi(tok.c_str()));
        }
        if (devs.empty()) {
            int n = 1;
            if (const char *env = std::getenv("HPFW_NUM_GPUS")) n = std::max(1, std::atoi(env));
            else n = std::max(1, hpfw_device_count());
            for (int d = 0; d < n; ++d) devs.push_back(d);
        }
        return devs;
    }
    hpfw_shard *require() const {
        if (!built) throw Error(HPFW_ERR_STATE, "ShardedMemoryStorage: build() or load() has not been called");
        return shard;
    }
    SearchResult to_result(const hpfw_match &m) const {
        if (m.track < 0) return {"", std::numeric_limits<size_t>::max(), 0};    // storage.h:28
        return {names[static_cast<size_t>(m.track)], static_cast<size_t>(m.cnt), m.offset};
    }
    std::vector<std::vector<SearchResult>> to_results(const std::vector<hpfw_match> &m, size_t nq, int topk) const {
        std::vector<std::vector<SearchResult>> out(nq);
        for (size_t q = 0; q < nq; ++q)
            for (int r = 0; r < topk; ++r) out[q].push_back(to_result(m[q * static_cast<size_t>(topk) + r]));
        return out;
    }
};

}  // namespace hpfw::db
