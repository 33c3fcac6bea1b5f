"""Parity of the CUDA matcher (hpfw_b200/csrc/matcher.cu, through the C ABI) with the oracle / the reference-generated
golden vectors. Integer work: every comparison is bit-exact."""
import numpy as np
import pytest

import oracle
from hpfw_b200 import MemoryStorage, synth
from hpfw_b200.api import SIZE_MAX

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[0, 1, 2, 3], ids=["popc", "tc_i8", "auto", "tc_fp4"], autouse=True)
def match_impl(ctx, request):
    """Every test of this file runs four times: integer-pipe kernel (matcher.cu), tensor-core kernel for every query
    (match_tc.cu) with int8 and with fp4 operands, and the default routing (tensor cores for well-filled groups of 128,
    integer pipes for the rest)."""
    from hpfw_b200._lib import check
    check(ctx._lib.hpfw_set_match_impl(ctx.handle, request.param))
    yield request.param
    check(ctx._lib.hpfw_set_match_impl(ctx.handle, 2))


def _triple(r):
    return (r.track, -1 if r.cnt >= (1 << 63) else r.cnt, r.offset)


def test_find_golden(ctx, matcher_golden):
    """MemoryStorage::find on every golden case (ragged, short refs, empty track, ties, empty DB, empty query)."""
    for name, c in matcher_golden.items():
        st = MemoryStorage(ctx).build_packed(c["words"], c["offs"])
        qo = c["qoffs"]
        for i in range(len(qo) - 1):
            got = _triple(st.find(c["qwords"][qo[i]:qo[i + 1]]))
            assert got == tuple(int(x) for x in c["res"][i]), (name, i)


def test_find_returns_filenames(ctx, collector_golden):
    c = collector_golden
    offs = c["offs"]
    pairs = [(str(n), c["words"][offs[i]:offs[i + 1]]) for i, n in enumerate(c["names"])]
    st = MemoryStorage(ctx).build(pairs)
    r = st.find(c["hpq"])
    assert (r.track, r.cnt, r.offset) == tuple(int(x) for x in c["res"])
    assert r.filename == str(c["names"][r.track])


def test_topk_batched_vs_oracle(ctx, matcher_golden):
    for name, c in matcher_golden.items():
        R = len(c["offs"]) - 1
        topk = R + 2
        st = MemoryStorage(ctx).build_packed(c["words"], c["offs"])
        out = st.find_topk_packed(c["qwords"], c["qoffs"], topk)
        qo = c["qoffs"]
        for i in range(len(qo) - 1):
            tr, d, o = oracle.find_topk(c["words"], c["offs"], c["qwords"][qo[i]:qo[i + 1]], topk)
            assert np.array_equal(out["track"][i], tr), (name, i)
            assert np.array_equal(out["cnt"][i], d), (name, i)
            assert np.array_equal(out["offset"][i], o), (name, i)


@pytest.mark.parametrize("seed,n_tracks,min_len,max_len,ks", [
    (1, 37, 0, 700, (1, 7, 8, 9, 63, 143, 385)),          # includes empty / shorter-than-query tracks
    (4, 37, 400, 700, (1, 7, 8, 9, 63, 143, 385)),
    (2, 5, 2100, 5000, (385, 1514, 2047, 2048, 2049)),
    (3, 64, 1, 2100, (16, 385)),
    (5, 16, 2040, 2060, (2, 3, 5, 385, 386, 387)),       # track lengths straddling the 2048-offset tile boundary
])
def test_random_ragged_vs_oracle(ctx, seed, n_tracks, min_len, max_len, ks):
    rng = np.random.default_rng(seed)
    lens = rng.integers(min_len, max_len, size=n_tracks)
    lens[rng.integers(0, n_tracks)] = max_len          # at least one long track
    words, offs = synth.synth_hashprint_db(seed, n_tracks, lens)
    kk = np.array([ks[i % len(ks)] for i in range(3 * len(ks) + 1)], dtype=np.int64)
    kk = np.minimum(kk, max_len)
    qw, qo, truth = synth.synth_hashprint_queries(seed + 100, words, offs, len(kk), kk)
    st = MemoryStorage(ctx).build_packed(words, offs)
    topk = 5
    out = st.find_topk_packed(qw, qo, topk)
    tr, d, o = oracle.find_topk_batch(words, offs, qw, qo, topk, 8)
    assert np.array_equal(out["track"], tr)
    assert np.array_equal(out["cnt"], d)
    assert np.array_equal(out["offset"], o)
    # noisy sub-sequences at 25 % bit flips are found at their true position whenever the query is long enough
    # (only when no reference is shorter than the query: a truncated query can win with a smaller distance, storage.h:34-38)
    long_q = (kk >= 63) & (kk <= lens.min())
    assert np.array_equal(out["track"][long_q, 0], truth[long_q, 0])
    assert np.array_equal(out["offset"][long_q, 0], truth[long_q, 1])


def test_random_queries_unrelated_to_db(ctx):
    """Random (non-planted) queries: near-ties everywhere, the strict-'<' tie rules decide."""
    rng = np.random.default_rng(9)
    words, offs = synth.synth_hashprint_db(9, 20, rng.integers(1, 900, size=20))
    ks = np.array([1, 2, 3, 4, 5, 8, 16, 33], dtype=np.int64)
    qo = np.zeros(len(ks) + 1, dtype=np.int64)
    np.cumsum(ks, out=qo[1:])
    qw = rng.integers(0, 1 << 64, size=int(qo[-1]), dtype=np.uint64)
    st = MemoryStorage(ctx).build_packed(words, offs)
    out = st.find_topk_packed(qw, qo, 20)
    tr, d, o = oracle.find_topk_batch(words, offs, qw, qo, 20, 8)
    assert np.array_equal(out["track"], tr) and np.array_equal(out["cnt"], d) and np.array_equal(out["offset"], o)


def test_sharded_merge_equals_unsharded(ctx):
    """DB split into shards with track_base, per-shard top-k keys merged by the merge kernel == unsharded result
    (the single-process version of the multi-GPU path)."""
    import torch
    from hpfw_b200.api import decode_keys
    rng = np.random.default_rng(11)
    n_tracks, topk = 41, 7
    words, offs = synth.synth_hashprint_db(11, n_tracks, rng.integers(50, 1500, size=n_tracks))
    qw, qo, _ = synth.synth_hashprint_queries(12, words, offs, 23, 50)
    whole = MemoryStorage(ctx).build_packed(words, offs).find_topk_packed(qw, qo, topk)
    dq = torch.from_numpy(qw.view(np.int64)).cuda()
    n_shards = 4
    bounds = np.linspace(0, n_tracks, n_shards + 1).astype(int)
    keys = torch.empty((n_shards, 23, topk), dtype=torch.int64, device="cuda")
    shards = []
    s = torch.cuda.current_stream().cuda_stream
    from hpfw_b200.api import stream_arg
    for g in range(n_shards):
        a, b = bounds[g], bounds[g + 1]
        st = MemoryStorage(ctx).build_packed(words[offs[a]:offs[b]], offs[a:b + 1] - offs[a], track_base=int(a))
        st.match_device(dq.data_ptr(), qo, topk, keys[g].data_ptr(), s)
        shards.append(st)
    merged = torch.empty((23, topk), dtype=torch.int64, device="cuda")
    import ctypes as C
    from hpfw_b200._lib import check
    check(ctx._lib.hpfw_topk_merge_device(ctx.handle, C.c_void_p(keys.data_ptr()), n_shards, 23, topk,
                                          C.c_void_p(merged.data_ptr()), stream_arg(s)))
    torch.cuda.synchronize()
    got = decode_keys(merged.cpu().numpy().view(np.uint64))
    for f in ("track", "cnt", "offset"):
        assert np.array_equal(got[f], whole[f]), f


def test_full_size_properties(ctx):
    """BASELINE config sizes (1000 x 14,411-word DB, 385-word queries): too big for the scalar oracle on every query, so
    check (i) planted queries are found at their true (track, offset) with the exact planted distance, (ii) a sample
    against the oracle, (iii) idempotence."""
    words, offs = synth.synth_hashprint_db(21, 1000, 14411)
    qw, qo, truth = synth.synth_hashprint_queries(22, words, offs, 64, 385)
    st = MemoryStorage(ctx).build_packed(words, offs)
    out = st.find_topk_packed(qw, qo, 10)
    assert np.array_equal(out["track"][:, 0], truth[:, 0])
    assert np.array_equal(out["offset"][:, 0], truth[:, 1])
    for i in range(64):
        a = offs[truth[i, 0]] + truth[i, 1]
        planted = int(np.unpackbits((words[a:a + 385] ^ qw[qo[i]:qo[i + 1]]).view(np.uint8)).sum())
        assert int(out["cnt"][i, 0]) == planted
    tr, d, o = oracle.find_topk_batch(words, offs, qw[:qo[4]], qo[:5], 10, 8)
    assert np.array_equal(out["track"][:4], tr) and np.array_equal(out["cnt"][:4], d) and np.array_equal(out["offset"][:4], o)
    again = st.find_topk_packed(qw, qo, 10)
    assert np.array_equal(again, out)


def _popcount_rows(x):
    return np.unpackbits(np.ascontiguousarray(x).view(np.uint8)).reshape(len(x), -1).sum(axis=1)


def test_extreme_distances(ctx):
    """Accumulator range of the int8 GEMM: exact copies (distance 0, dot = +64 k), bitwise complements (distance 64 k,
    dot = -64 k), all-zero / all-one words, at the longest query the ABI accepts (4096 words) and at odd lengths."""
    rng = np.random.default_rng(41)
    lens = np.array([4096, 5000, 4609, 700, 4097], dtype=np.int64)
    words, offs = synth.synth_hashprint_db(41, len(lens), lens)
    words = words.copy()
    words[offs[3]:offs[3] + 300] = 0
    words[offs[3] + 300:offs[3] + 600] = np.uint64(0xFFFFFFFFFFFFFFFF)
    queries = []
    for k in (4096, 4095, 385, 193, 192, 5, 1):
        a = int(offs[1]) + 17
        queries.append(words[a:a + k].copy())                       # exact copy of track 1 at offset 17
        queries.append(~words[a:a + k])                             # its complement
    # a long exact run followed (or preceded) by unrelated words: large partial sums, then many small increments
    half = words[int(offs[1]) + 100:int(offs[1]) + 100 + 2048]
    noise = rng.integers(0, 1 << 64, size=2048, dtype=np.uint64)
    queries.append(np.concatenate([half, noise]))
    queries.append(np.concatenate([noise, words[int(offs[1]) + 2148:int(offs[1]) + 4196]]))
    queries.append(np.concatenate([~half, noise]))
    queries.append(np.zeros(300, dtype=np.uint64))
    queries.append(np.full(300, 0xFFFFFFFFFFFFFFFF, dtype=np.uint64))
    queries.append(np.zeros(0, dtype=np.uint64))                    # empty query: distance 0 at offset 0 of track 0
    qo = np.zeros(len(queries) + 1, dtype=np.int64)
    np.cumsum([len(q) for q in queries], out=qo[1:])
    qw = np.concatenate(queries)
    st = MemoryStorage(ctx).build_packed(words, offs)
    out = st.find_topk_packed(qw, qo, len(lens))
    tr, d, o = oracle.find_topk_batch(words, offs, qw, qo, len(lens), 8)
    assert np.array_equal(out["track"], tr) and np.array_equal(out["cnt"], d) and np.array_equal(out["offset"], o)
    assert out["cnt"][0, 0] == 0 and out["track"][0, 0] == 1 and out["offset"][0, 0] == 17


def test_planted_at_every_column_range(ctx):
    """Exact copies planted around every boundary of the tensor-core tiles (480 / 512 offsets per tile, two N halves of
    240 / 256, 64-column TMEM loads, the fp4 kernel's scale-factor columns 480..511) in a group that is mostly padding."""
    rng = np.random.default_rng(61)
    n, k = 2200, 50
    words = rng.integers(0, 1 << 64, size=n, dtype=np.uint64)
    offs = np.array([0, n], dtype=np.int64)
    plant = [0, 5, 63, 64, 223, 224, 239, 240, 255, 256, 300, 447, 448, 470, 479, 480, 511, 512, 600, 959, 960, 1023, 1024,
             1100, 1439, 1440, 1535, 1536, 2047, 2048, n - k]
    qw = np.concatenate([words[p:p + k] for p in plant])
    qo = np.arange(len(plant) + 1, dtype=np.int64) * k
    out = MemoryStorage(ctx).build_packed(words, offs).find_topk_packed(qw, qo, 1)
    assert np.array_equal(out["offset"][:, 0], np.array(plant))
    assert not out["cnt"][:, 0].any()


def test_many_queries_mixed_lengths(ctx):
    """300 queries of 7 different lengths against ragged tracks: several groups of 128 with different k inside one group
    (zero-padded query rows), a partial last group, queries longer than some tracks."""
    rng = np.random.default_rng(51)
    n_tracks = 24
    lens = rng.integers(100, 1800, size=n_tracks)
    lens[3] = 30
    lens[7] = 513
    lens[8] = 512
    lens[9] = 511
    words, offs = synth.synth_hashprint_db(51, n_tracks, lens)
    ks = np.array([63, 64, 65, 143, 385, 386, 40], dtype=np.int64)
    kk = ks[rng.integers(0, len(ks), size=300)]
    qw, qo, _ = synth.synth_hashprint_queries(52, words, offs, len(kk), kk)
    st = MemoryStorage(ctx).build_packed(words, offs)
    out = st.find_topk_packed(qw, qo, 6)
    tr, d, o = oracle.find_topk_batch(words, offs, qw, qo, 6, 8)
    assert np.array_equal(out["track"], tr) and np.array_equal(out["cnt"], d) and np.array_equal(out["offset"], o)


def test_query_chunks(ctx, monkeypatch):
    """More queries than the best[] scratch holds at once (forced here with a 64 KB bound instead of the 1 GiB default):
    the batch is processed in chunks, each with its own routing, group tables and key slice."""
    monkeypatch.setenv("HPFW_MATCH_SCRATCH_BYTES", str(64 * 1024))
    rng = np.random.default_rng(71)
    n_tracks = 37
    words, offs = synth.synth_hashprint_db(71, n_tracks, rng.integers(60, 900, size=n_tracks))
    kk = np.array([50, 51, 143, 40])[rng.integers(0, 4, size=700)]          # 64 KB / (37 * 8 B) = 221 queries per chunk
    qw, qo, truth = synth.synth_hashprint_queries(72, words, offs, len(kk), kk)
    out = MemoryStorage(ctx).build_packed(words, offs).find_topk_packed(qw, qo, 4)
    tr, d, o = oracle.find_topk_batch(words, offs, qw, qo, 4, 8)
    assert np.array_equal(out["track"], tr) and np.array_equal(out["cnt"], d) and np.array_equal(out["offset"], o)


def test_limits_reported(ctx):
    from hpfw_b200 import HpfwError
    words, offs = synth.synth_hashprint_db(31, 2, 100)
    st = MemoryStorage(ctx).build_packed(words, offs)
    with pytest.raises(HpfwError):
        st.find(np.zeros(5000, dtype=np.uint64))      # > HPFW_MAX_QUERY_WORDS
