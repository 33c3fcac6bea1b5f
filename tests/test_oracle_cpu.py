"""The oracle (oracle/hpfw_oracle.c) pinned against golden vectors produced by the reference's own headers
(tests/golden/make_golden.py -> oracle/_ref), and — when oracle/_ref is present on this host — against that library live."""
import numpy as np
import pytest

import oracle

SIZE_MAX = (1 << 64) - 1


def _res(t):
    tr, cnt, off = t
    return (tr, -1 if cnt >= (1 << 63) else cnt, off)


def test_find_matches_reference_golden(matcher_golden):
    for name, c in matcher_golden.items():
        qo = c["qoffs"]
        for i in range(len(qo) - 1):
            got = _res(oracle.find(c["words"], c["offs"], c["qwords"][qo[i]:qo[i + 1]]))
            assert got == tuple(int(x) for x in c["res"][i]), (name, i)


def test_find_edge_semantics(matcher_golden):
    # storage.h:28 — empty DB leaves {"" , SIZE_MAX, 0}
    c = matcher_golden["empty_db"]
    assert tuple(c["res"][0]) == (-1, -1, 0)
    # an empty reference track gives distance 0 and wins (k = min(k, 0) = 0, one offset)
    c = matcher_golden["empty_track"]
    assert all(int(r[0]) == 1 and int(r[1]) == 0 and int(r[2]) == 0 for r in c["res"])
    # ties: earliest track, lowest offset
    c = matcher_golden["ties"]
    assert tuple(c["res"][0]) == (0, 0, 30)      # base[30:80] is also in track 2: track 0 wins
    assert tuple(c["res"][1]) == (1, 0, 5)       # periodic: offsets 5, 30, 55, ... all exact: lowest wins
    assert tuple(c["res"][2]) == (0, 0, 0)


def test_topk_consistent_with_find(matcher_golden):
    for name, c in matcher_golden.items():
        if name == "empty_db":
            continue
        qo = c["qoffs"]
        R = len(c["offs"]) - 1
        for i in range(len(qo) - 1):
            q = c["qwords"][qo[i]:qo[i + 1]]
            tr, d, o = oracle.find_topk(c["words"], c["offs"], q, R + 2)
            assert (int(tr[0]), int(d[0]), int(o[0])) == tuple(int(x) for x in c["res"][i])
            assert list(tr[R:]) == [-1, -1]
            dist, off = oracle.per_track_best(c["words"], c["offs"], q)
            order = sorted(range(R), key=lambda r: (int(dist[r]), r))
            assert list(tr[:R]) == order
            assert [int(x) for x in d[:R]] == [int(dist[r]) for r in order]
            assert [int(x) for x in o[:R]] == [int(off[r]) for r in order]


def test_batch_equals_single(matcher_golden):
    c = matcher_golden["ragged"]
    tr, d, o = oracle.find_topk_batch(c["words"], c["offs"], c["qwords"], c["qoffs"], 3, 4)
    for i in range(len(c["qoffs"]) - 1):
        t1, d1, o1 = oracle.find_topk(c["words"], c["offs"], c["qwords"][c["qoffs"][i]:c["qoffs"][i + 1]], 3)
        assert np.array_equal(tr[i], t1) and np.array_equal(d[i], d1) and np.array_equal(o[i], o1)


def test_calc_frames_kat(kat_golden):
    S = kat_golden["frames_in"]
    fr = np.zeros(2420 * 11, dtype=np.float32)
    n = oracle.lib().orc_calc_frames(np.ascontiguousarray(S).reshape(-1), 30, fr)
    assert n == 11
    fr = fr.reshape(2420, 11)
    assert np.array_equal(fr, kat_golden["frames_out"])
    # band-major, context inner (SURVEY Appendix A probe 1)
    assert list(fr[:4, 0]) == [0, 1, 2, 3] and fr[20, 0] == 100 and fr[21, 3] == 104


def test_bit_order_kat(kat_golden):
    for f in (0, 1, 31, 62, 63):
        y = np.zeros((81, 64), dtype=np.float32)
        y[0, :] = -1.0
        y[0, f] = 1.0
        hp = np.zeros(1, dtype=np.uint64)
        assert oracle.lib().orc_fingerprint_pack_f32(y.reshape(-1), 81, hp) == 1
        assert hp[0] == kat_golden[f"only_{f}"][0] == np.uint64(1) << np.uint64(63 - f)
    hp = np.zeros(1, dtype=np.uint64)
    oracle.lib().orc_fingerprint_pack_f32(np.zeros(81 * 64, dtype=np.float32), 81, hp)
    assert hp[0] == kat_golden["all_zero_delta"][0] == np.uint64(SIZE_MAX)   # delta == 0 -> bit 1 (hashprint_handle.h:121)


def test_amplitude_to_db_kat(kat_golden):
    s = np.ascontiguousarray(kat_golden["amp"].copy()).reshape(-1)
    oracle.lib().orc_amplitude_to_db(s, 50)
    # the reference evaluates log10 in float under -ffast-math; tolerance 1e-4 dB (values span [-80, 0])
    assert np.max(np.abs(s.reshape(50, 121) - kat_golden["db"])) < 1e-4
    assert s.max() == 0.0 and s.min() == -80.0


def test_projection_and_hashprint_vs_reference_golden(hashprint_golden):
    g = hashprint_golden
    spec, filt = g["spec0"], g["filters"]
    cols = spec.shape[0]
    y = np.zeros((cols - 19) * 64, dtype=np.float32)
    oracle.lib().orc_project_f32(spec.reshape(-1), cols, filt.reshape(-1), y)
    yh = y.reshape(-1, 64)[:256]
    scale = np.abs(g["y0_head"]).max()
    # fp32 summation order differs between Eigen's GEBP and a sequential loop: relative tolerance 2e-5 of max|y|
    assert np.max(np.abs(yh - g["y0_head"])) <= 2e-5 * scale
    for key_s, key_h in (("spec0", "hp0"), ("q_spec", "hpq")):
        hp = oracle.hashprint_from_spectrogram(g[key_s], filt)
        ref = g[key_h]
        assert hp.shape == ref.shape
        diff = int(np.unpackbits((hp ^ ref).view(np.uint8)).sum())
        assert diff <= 1e-3 * 64 * len(ref), (key_s, diff)       # >= 99.9 % of bits (north_star)
        # the differing bits are exactly the near-zero deltas: compare with the rounding-free yardstick
        hp64, margin = oracle.hashprint_f64(g[key_s], filt)
        bad = np.unpackbits((hp64 ^ ref).view(np.uint8)).sum()
        assert bad <= 1e-3 * 64 * len(ref)


def test_amplitude_to_db_on_cqt_magnitudes(hashprint_golden):
    g = hashprint_golden
    s = np.ascontiguousarray(g["mag0"].copy()).reshape(-1)
    oracle.lib().orc_amplitude_to_db(s, g["mag0"].shape[0])
    assert np.max(np.abs(s.reshape(g["spec0"].shape) - g["spec0"])) < 1e-3


def test_collector_end_to_end_golden(collector_golden, hashprint_golden):
    c = collector_golden
    got = _res(oracle.find(c["words"], c["offs"], c["hpq"]))
    assert got == tuple(int(x) for x in c["res"])


def test_nsgcq_design_sizes():
    from oracle import nsgcq
    # SURVEY.md §8 derived sizes
    for n, m, cols in ((7938000, 43528, 14510), (1323000, 7255, 2419), (661500, 3627, 1210), (882000, 4836, 1613),
                       (264600, 1451, 484), (132300, 725, 242), (88200, 484, 162)):
        pos, lg, mm = nsgcq.nsg_design(n)
        assert mm == m and nsgcq.spectrogram_cols(n) == cols
        assert lg.min() >= 96 and len(pos) == 121


def test_nsgcq_reproduces_golden_magnitudes(hashprint_golden):
    from oracle import nsgcq
    g = hashprint_golden
    mag = nsgcq.nsgcq_magnitude(g["audio_q"]).astype(np.float32)
    assert np.array_equal(mag, g["q_mag"])


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built on this host")
def test_live_reference_agrees_on_random_cases():
    from hpfw_b200 import synth
    rng = np.random.default_rng(5)
    for trial in range(5):
        lens = rng.integers(0, 400, size=6)
        w, o = synth.synth_hashprint_db(200 + trial, len(lens), lens)
        for k in (1, 17, 64, 150, 450):
            q = rng.integers(0, 1 << 64, size=k, dtype=np.uint64)
            assert _res(oracle.find(w, o, q)) == _res(oracle.ref_find(w, o, q))


def test_nsgcq_window_variants():
    """The "hann" convention is a switch (oracle/nsgcq.py: band_window): the periodic form peaks at exactly 1 on the band
    centre; the symmetric form is 0 on both end taps and mirror-symmetric over them. They differ by O(1/L) per tap."""
    from oracle import nsgcq
    for L in (96, 97, 725, 4836):
        k, wp = nsgcq.band_window(L, "periodic")
        k2, ws = nsgcq.band_window(L, "symmetric")
        assert np.array_equal(k, k2) and len(k) == L and k[0] == -(L // 2)
        assert wp[L // 2] == 1.0 and np.all(wp <= 1.0)
        assert abs(ws[0]) < 1e-12 and abs(ws[-1]) < 1e-12 and np.allclose(ws, ws[::-1], atol=1e-12)
        assert np.max(np.abs(wp - ws)) > 0.5 / L
    with pytest.raises(ValueError):
        nsgcq.band_window(96, "hamming")
    x = np.random.default_rng(0).standard_normal(88200).astype(np.float32)
    a, b = nsgcq.nsgcq_magnitude(x, "periodic"), nsgcq.nsgcq_magnitude(x, "symmetric")
    assert a.shape == b.shape and np.max(np.abs(a - b)) > 1e-4 * a.max()
