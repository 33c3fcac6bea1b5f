// include/hpfw/audioproblems/live-song-id/storage.h — db::MemoryStorage on the GPU.
//
// Same interface and result semantics as /root/reference/include/hpfw/audioproblems/live-song-id/storage.h:8-93:
// build(pairs), find(hashprint) -> SearchResult{filename, cnt, offset}, save/load of the cereal dump. The DB lives in HBM;
// find() runs the exhaustive Hamming cross-correlation kernel (matcher.cu). Additions: find_topk (the notebook's ranking,
// examples/python/liveid.ipynb:909-927) and batched queries.
#pragma once

#include <cstdint>
#include <functional>
#include <limits>
#include <memory>
#include <optional>
#include <string>
#include <utility>
#include <vector>

#include "../../device.h"
#include "../../io/cereal_compat.h"

namespace hpfw::db {

template <typename Collector>
class MemoryStorage {
public:
    using Hashprint = typename Collector::Hashprint;
    using Pair = typename Collector::FilenameFingerprintPair;

    struct SearchResult {        // storage.h:11-15
        std::string filename;
        size_t cnt;
        int64_t offset;
    };

    explicit MemoryStorage(int device = 0) : ctx(device::Context::shared(device)) {}
    ~MemoryStorage() { hpfw_db_destroy(db); }
    MemoryStorage(const MemoryStorage &) = delete;
    MemoryStorage &operator=(const MemoryStorage &) = delete;

    /// DB order = order of `fingerprints` (the reference moves the collector's vector in, storage.h:21-25).
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
        upload(words, offs);
        pending_words = nullptr;
        host_words = std::move(words);
        host_offs = std::move(offs);
    }

    /// build() from hashprints that are still in HBM (Collector::prepare_device): device-to-device, the words never visit
    /// the host. DB order = order of d.names. The host copy save() needs is fetched lazily.
    template <typename DeviceHashprints>
    void build_device(const DeviceHashprints &d) {
        names = d.names;
        host_words.clear();
        host_offs.assign(1, 0);
        for (int w : d.words) host_offs.push_back(host_offs.back() + w);
        std::scoped_lock l(ctx->mutex());
        hpfw_db_destroy(db);
        db = nullptr;
        device::check(hpfw_xs_build_db(d.xs->get(), d.order.data(), static_cast<int>(d.order.size()), 0, &db));
        // the dump written by save() needs the words on the host: one contiguous copy per track, only when asked for
        pending_words = [xs = d.xs, gen = d.generation, order = d.order, words = d.words, ctx = ctx]() {
            if (xs->generation != gen)
                throw Error(HPFW_ERR_STATE, "MemoryStorage::save: the collector has started another batch since build_device(); "
                                            "the hashprints are no longer resident in its stream (save() right after index())");
            std::vector<uint64_t> out;
            size_t total = 0;
            for (int w : words) total += static_cast<size_t>(w);
            out.resize(total);
            size_t pos = 0;
            std::scoped_lock l2(ctx->mutex());
            for (size_t i = 0; i < order.size(); ++i) {
                device::check(hpfw_xs_hashprint_host(xs->get(), order[i], out.data() + pos));
                pos += static_cast<size_t>(words[i]);
            }
            return out;
        };
    }

    /// Batched top-k of queries whose hashprints are still in HBM (Collector::calc_hashprints_device), one result list per
    /// query in d.names order.
    template <typename DeviceHashprints>
    std::vector<std::vector<SearchResult>> find_topk_device(const DeviceHashprints &d, int topk) const {
        const size_t nq = d.order.size();
        std::vector<std::vector<SearchResult>> out(nq);
        if (nq == 0) return out;
        std::vector<hpfw_match> m(nq * static_cast<size_t>(topk));
        {
            std::scoped_lock l(ctx->mutex());
            device::check(hpfw_xs_match(d.xs->get(), require(), topk, m.data()));
        }
        // the stream holds exactly the tracks of d.order, ascending (calc_hashprints_device sorts by stream track)
        for (size_t q = 0; q < nq; ++q)
            for (int r = 0; r < topk; ++r) out[q].push_back(to_result(m[q * topk + r]));
        return out;
    }

    /// MemoryStorage::find (storage.h:27-64): best track by strict '<' over (distance, then DB order), lowest offset.
    SearchResult find(const Hashprint &hp) const {
        hpfw_match m;
        std::scoped_lock l(ctx->mutex());
        device::check(hpfw_db_find(require(), hp.data(), static_cast<int>(hp.size()), &m));
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
        {
            std::scoped_lock l(ctx->mutex());
            device::check(hpfw_db_find_topk(require(), qw.data(), qo.data(), static_cast<int>(queries.size()), topk,
                                            m.data()));
        }
        std::vector<std::vector<SearchResult>> out(queries.size());
        for (size_t q = 0; q < queries.size(); ++q)
            for (int r = 0; r < topk; ++r) out[q].push_back(to_result(m[q * topk + r]));
        return out;
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
            v.emplace_back(names[r], std::vector<uint64_t>(host_words.begin() + host_offs[r],
                                                           host_words.begin() + host_offs[r + 1]));
        io::save_db(dump_name, v);
        return dump_name;
    }

    MemoryStorage &load(const std::string &dump_name) {   // storage.h:79-86
        std::vector<Pair> pairs;
        for (auto &e : io::load_db(dump_name)) pairs.push_back({std::move(e.first), std::move(e.second)});
        build(std::move(pairs));
        return *this;
    }

    size_t size() const { return names.size(); }

private:
    std::shared_ptr<device::Context> ctx;
    hpfw_db *db = nullptr;
    std::vector<std::string> names;
    mutable std::vector<uint64_t> host_words;     // kept for save(); the matcher only reads the HBM copy
    std::vector<int64_t> host_offs;
    mutable std::function<std::vector<uint64_t>()> pending_words;   // build_device: host copy fetched on the first save()

    hpfw_db *require() const {
        if (!db) throw Error(HPFW_ERR_STATE, "MemoryStorage: build() or load() has not been called");
        return db;
    }
    void upload(const std::vector<uint64_t> &words, const std::vector<int64_t> &offs) {
        std::scoped_lock l(ctx->mutex());
        hpfw_db_destroy(db);
        db = nullptr;
        device::check(hpfw_db_build(ctx->get(), words.data(), offs.data(), static_cast<int>(offs.size()) - 1, 0, &db));
    }
    SearchResult to_result(const hpfw_match &m) const {
        if (m.track < 0) return {"", std::numeric_limits<size_t>::max(), 0};    // storage.h:28
        return {names[static_cast<size_t>(m.track)], static_cast<size_t>(m.cnt), m.offset};
    }
};

}  // namespace hpfw::db
