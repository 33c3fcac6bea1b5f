"""Stage 1 on the GPU: the FFT engine against numpy, the CQT against the oracle restatement (oracle/nsgcq.py; "parity
unpinned" versus essentia, see its header), and the audio -> hashprint path against the reference-generated goldens.

Tolerances (north_star: "within a stated fp32 relative tolerance"):
  FFT:        max |err| <= 2e-6 * max |X|      (fp32 Stockham, up to 4M points)
  magnitude:  max |err| <= 1e-5 * max |mag|    per track
  dB:         <= 0.01 dB wherever the oracle is above -79 dB (Bluestein path for non-smooth lengths and tracks longer
              than 6.7 minutes: 0.01 dB above -70 dB, 0.05 dB above -79 dB)
"""
import ctypes as C

import numpy as np
import pytest

from hpfw_b200 import synth
from hpfw_b200._lib import check
from oracle import nsgcq

pytestmark = pytest.mark.gpu


def _fft(ctx, x, inverse=False):
    x = np.ascontiguousarray(x, dtype=np.complex64)
    out = np.empty_like(x)
    check(ctx._lib.hpfw_fft_c2c(ctx.handle, x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), len(x),
                                1 if inverse else 0))
    return out


@pytest.mark.parametrize("n", [64, 96, 210, 1024, 4410, 6300, 65536, 66150, 132300, 3969000,
                               1 << 22])
def test_fft_engine_vs_numpy(ctx, n):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    ref = np.fft.fft(x.astype(np.complex128))
    got = _fft(ctx, x)
    assert np.max(np.abs(got - ref)) <= 2e-6 * np.max(np.abs(ref))
    refi = np.fft.ifft(x.astype(np.complex128)) * n
    goti = _fft(ctx, x, inverse=True)
    assert np.max(np.abs(goti - refi)) <= 2e-6 * np.max(np.abs(refi))


def test_fft_rejects_non_smooth_length(ctx):
    from hpfw_b200 import HpfwError
    from hpfw_b200._lib import ERR_LIMIT
    with pytest.raises(HpfwError) as e:
        _fft(ctx, np.zeros(2 * 5441, dtype=np.complex64))
    assert e.value.code == ERR_LIMIT


def _cqt(ctx, audio, magnitude=False):
    audio = np.ascontiguousarray(audio, dtype=np.float32)
    cols = ctx._lib.hpfw_cqt_cols(len(audio))
    out = np.zeros((cols, 121), dtype=np.float32)
    got = C.c_int()
    fn = ctx._lib.hpfw_cqt_magnitude if magnitude else ctx._lib.hpfw_cqt_spectrogram
    check(fn(ctx.handle, audio.ctypes.data_as(C.c_void_p), len(audio), out.ctypes.data_as(C.c_void_p), C.byref(got)))
    assert got.value == cols
    return out


@pytest.mark.parametrize("seconds,sr", [(6.0, 22050), (6.0, 44100), (30.0, 22050), (20.0, 44100), (2.0, 44100),
                                        (180.0, 44100),
                                        (450.0, 44100), (780.0, 44100),     # > 6.7 min: chirp-z column length 32
                                        (1500.0, 44100)])                   # > 13.5 min: column length 64
def test_cqt_vs_oracle(ctx, seconds, sr):
    audio = synth.synth_track(int(seconds * 10) + sr, seconds, sr)
    ref_mag = nsgcq.nsgcq_magnitude(audio)
    mag = _cqt(ctx, audio, magnitude=True)
    assert mag.shape == ref_mag.shape
    scale = np.abs(ref_mag).max()
    assert np.max(np.abs(mag - ref_mag)) <= 1e-5 * scale
    ref_db = nsgcq.amplitude_to_db(ref_mag)
    db = _cqt(ctx, audio)
    above = ref_db > -79.0
    assert above.mean() > 0.5
    if seconds <= 400.0:
        assert np.max(np.abs(db[above] - ref_db[above])) <= 0.01
    else:
        # 20-35 M-point transforms: the fp32 rounding noise (~3e-7 of the track maximum) is 0.2 % of an amplitude that sits
        # 79 dB below it; same bar as the Bluestein path
        assert np.max(np.abs(db[ref_db > -70.0] - ref_db[ref_db > -70.0])) <= 0.01
        assert np.max(np.abs(db[above] - ref_db[above])) <= 0.05
    assert db.max() == 0.0 and db.min() >= -80.0
    m = nsgcq.nsg_design(len(audio))[2]
    if m % 3 == 0:      # the column the reference leaves unwritten (cqt.h:73-81): defined as amplitude 0 -> the floor
        assert np.all(db[-1] == -80.0) and np.all(mag[-1] == 0.0)


def test_cqt_golden_query(ctx, hashprint_golden):
    g = hashprint_golden
    mag = _cqt(ctx, g["audio_q"], magnitude=True)
    assert np.max(np.abs(mag - g["q_mag"])) <= 1e-5 * np.abs(g["q_mag"]).max()
    db = _cqt(ctx, g["audio_q"])
    above = g["q_spec"] > -79.0
    assert np.max(np.abs(db[above] - g["q_spec"][above])) <= 0.01


def test_audio_to_hashprint_vs_reference_golden(ctx, hashprint_golden, collector_golden):
    """a8 = ParallelCollector::calc_hashprint on a decoded buffer: GPU CQT + projection vs (oracle CQT -> reference headers).
    >= 99.9 % of the bits (north_star) and the identical top-1 match."""
    from hpfw_b200 import MemoryStorage
    import oracle
    g, c = hashprint_golden, collector_golden
    f = np.ascontiguousarray(g["filters"])
    check(ctx._lib.hpfw_set_filters(ctx.handle, f.ctypes.data_as(C.c_void_p)))
    audio = np.ascontiguousarray(g["audio_q"])
    n_exp = ctx._lib.hpfw_hashprint_words_for_samples(len(audio))
    assert n_exp == len(g["hpq"])
    hp = np.zeros(n_exp, dtype=np.uint64)
    n = C.c_int()
    check(ctx._lib.hpfw_calc_hashprint_audio(ctx.handle, audio.ctypes.data_as(C.c_void_p), len(audio),
                                             hp.ctypes.data_as(C.c_void_p), C.byref(n)))
    assert n.value == n_exp
    diff = int(np.unpackbits((hp ^ g["hpq"]).view(np.uint8)).sum())
    assert diff <= 1e-3 * 64 * n_exp, f"{diff} of {64 * n_exp} bits differ"
    st = MemoryStorage(ctx).build_packed(c["words"], c["offs"])
    r = st.find(hp)
    ref = oracle.find(c["words"], c["offs"], g["hpq"])
    assert (r.track, r.offset) == (ref[0], ref[2])


@pytest.mark.parametrize("n", [264601,          # odd
                               2 * 131071,      # N/2 prime: no smooth split
                               661501,          # 30 s @22.05 kHz + 1 sample, prime-ish
                               132300 + 2])     # even, N/2 = 66151 = 83 * 797
def test_cqt_any_length_vs_oracle(ctx, n):
    """Lengths whose half is not {2,3,5,7}-smooth take the Bluestein (chirp-convolution) path: same tolerances."""
    sr = 44100
    audio = synth.synth_track(n % 1000, n / sr + 0.01, sr)[:n]
    assert len(audio) == n
    ref_mag = nsgcq.nsgcq_magnitude(audio)
    mag = _cqt(ctx, audio, magnitude=True)
    assert mag.shape == ref_mag.shape
    assert np.max(np.abs(mag - ref_mag)) <= 1e-5 * np.abs(ref_mag).max()
    ref_db = nsgcq.amplitude_to_db(ref_mag)
    db = _cqt(ctx, audio)
    # the chirp convolution is ~2x longer than the packed FFT: absolute error ~1.5e-7 of the maximum, which is
    # 0.01 dB at -70 dB and 0.03 dB at -79 dB
    assert np.max(np.abs(db[ref_db > -70.0] - ref_db[ref_db > -70.0])) <= 0.01
    assert np.max(np.abs(db[ref_db > -79.0] - ref_db[ref_db > -79.0])) <= 0.05


def test_cqt_bluestein_matches_packed_path(ctx, monkeypatch):
    """The same smooth-length audio through both FFT routes gives the same magnitudes (to fp32 rounding)."""
    audio = synth.synth_track(5, 6.0, 44100)
    a = _cqt(ctx, audio, magnitude=True)
    import hpfw_b200
    monkeypatch.setenv("HPFW_CQT_BLUESTEIN", "1")
    ctx2 = hpfw_b200.Context(0)        # plans are cached per context: a fresh one re-plans under the override
    try:
        b = _cqt(ctx2, audio, magnitude=True)
    finally:
        ctx2.close()
    assert np.max(np.abs(a - b)) <= 2e-6 * np.abs(a).max()


def test_batch_of_distinct_lengths_equals_single_calls(ctx, hashprint_golden):
    """A library whose tracks all have different sample counts: every track plans (more lengths than the plan cache
    holds; smooth, non-smooth and odd lengths mixed; four lanes). The batched device entry must give exactly the
    per-track result of a fresh context."""
    import torch
    import hpfw_b200
    ex = hpfw_b200.HashprintExtractor(ctx)
    ex.set_filters(hashprint_golden["filters"])
    base = synth.synth_track(77, 4.0, 44100)
    # even lengths first (the packed path wants 8-byte aligned tracks), odd lengths (Bluestein) last
    lens = [88200 + 882 * i for i in range(40)] + [90002 + 2 * i for i in range(8)] + [88201 + 2 * i for i in range(8)]
    parts = [base[:n] * np.float32(0.5 + 0.01 * i) for i, n in enumerate(lens)]
    so = np.zeros(len(lens) + 1, dtype=np.int64)
    so[1:] = np.cumsum(lens)
    words = [ex.words(n) for n in lens]
    d_audio = torch.from_numpy(np.concatenate(parts)).cuda()
    d_hp = torch.zeros(int(sum(words)), dtype=torch.int64, device="cuda")
    ex.calc_hashprint_batch_device(d_audio.data_ptr(), so, d_hp.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = d_hp.cpu().numpy().view(np.uint64)
    ctx2 = hpfw_b200.Context(0)
    try:
        ex2 = hpfw_b200.HashprintExtractor(ctx2)
        ex2.set_filters(hashprint_golden["filters"])
        w0 = 0
        for n, part, w in zip(lens, parts, words):
            ref = ex2.calc_hashprint(part)
            assert len(ref) == w
            assert np.array_equal(got[w0:w0 + w], ref), f"track of {n} samples differs"
            w0 += w
    finally:
        ctx2.close()


def test_batch_of_repeated_lengths_graph_replay(ctx, hashprint_golden):
    """Tracks of equal length go four at a time through a captured launch graph (cqt.cu cqt_gang_run) whose input / output
    nodes are re-pointed on every replay; lengths that occur fewer than 8 times, and odd (Bluestein) lengths, take the
    kernel-by-kernel path in the same call. Two calls with the audio at different device addresses (the second replays the
    first's graphs), then one with a longer track first (the lanes' scratch grows: the graphs are re-captured): every
    hashprint must equal the single-track call of a fresh context."""
    import torch
    import hpfw_b200
    ex = hpfw_b200.HashprintExtractor(ctx)
    ex.set_filters(hashprint_golden["filters"])
    base = synth.synth_track(78, 9.0, 44100)
    ctx2 = hpfw_b200.Context(0)
    try:
        ex2 = hpfw_b200.HashprintExtractor(ctx2)
        ex2.set_filters(hashprint_golden["filters"])
        for rep, lens in enumerate([[132300] * 19 + [176400] * 9 + [154350] * 5 + [132301] * 2,
                                    [132300] * 11 + [176400] * 8,
                                    [352800] + [132300] * 9 + [176400] * 8]):
            rng = np.random.default_rng(100 + rep)
            parts = []
            for i, n in enumerate(lens):
                o = int(rng.integers(0, len(base) - n)) if n < len(base) else 0
                seg = base[o:o + n] if n <= len(base) else np.concatenate([base, base])[:n]
                parts.append((seg * np.float32(0.4 + 0.02 * i)).astype(np.float32))
            so = np.zeros(len(lens) + 1, dtype=np.int64)
            so[1:] = np.cumsum(lens)
            # an odd length would misalign the tracks behind it: pad every track start to an even sample offset
            starts, total = [], 0
            for n in lens:
                starts.append(total)
                total += n + (n & 1)
            pad = rep * 1024                      # a different base address on every call
            host = np.zeros(total + pad, dtype=np.float32)
            for st, part in zip(starts, parts):
                host[pad + st:pad + st + len(part)] = part
            words = [ex.words(n) for n in lens]
            d_all = torch.from_numpy(host).cuda()
            d_hp = torch.zeros(int(sum(words)), dtype=torch.int64, device="cuda")
            if all(n % 2 == 0 for n in lens):
                ex.calc_hashprint_batch_device(d_all.data_ptr() + 4 * pad, so, d_hp.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream)
            else:
                # contiguous offsets are the ABI: run the even-length prefix and the odd tail as two calls
                k = next(i for i, n in enumerate(lens) if n % 2)
                ex.calc_hashprint_batch_device(d_all.data_ptr() + 4 * pad, so[:k + 1], d_hp.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream)
                w0 = int(sum(words[:k]))
                for i in range(k, len(lens)):
                    one = np.array([0, lens[i]], dtype=np.int64)
                    ex.calc_hashprint_batch_device(d_all.data_ptr() + 4 * (pad + starts[i]), one, d_hp.data_ptr() + 8 * w0,
                                                   torch.cuda.current_stream().cuda_stream)
                    w0 += words[i]
            torch.cuda.synchronize()
            got = d_hp.cpu().numpy().view(np.uint64)
            w0 = 0
            for i, (n, part, w) in enumerate(zip(lens, parts, words)):
                ref = ex2.calc_hashprint(part)
                assert np.array_equal(got[w0:w0 + w], ref), f"call {rep}, track {i} ({n} samples) differs"
                w0 += w
    finally:
        ctx2.close()


def test_cqt_limits(ctx):
    from hpfw_b200 import HpfwError
    from hpfw_b200._lib import ERR_SHORT
    with pytest.raises(HpfwError) as e:
        _cqt(ctx, np.zeros(4000, dtype=np.float32))          # too short for the 121-band design
    assert e.value.code == ERR_SHORT
    from hpfw_b200._lib import ERR_LIMIT
    with pytest.raises(HpfwError) as e:
        _cqt(ctx, np.zeros(44100 * 1800, dtype=np.float32))   # 30 min: beyond the 64 x 8192-point chirp-z transform
    assert e.value.code == ERR_LIMIT and "27 minutes" in str(e.value)


def test_pcm16_entry_points_equal_the_float_path(ctx, hashprint_golden):
    """16-bit PCM in (host, device, batch): the device does MonoLoader's sample / 32768 (exact in float), so the hashprints
    are those of the float entry points on the converted samples, bit for bit; odd lengths and unaligned batch offsets too."""
    import torch
    import hpfw_b200
    ex = hpfw_b200.HashprintExtractor(ctx)
    ex.set_filters(np.ascontiguousarray(hashprint_golden["filters"]))
    lens = [88200, 132300, 100001, 264600]
    pcm = [np.clip(np.round(synth.synth_track(70 + i, n / 44100.0 + 0.01, 44100)[:n] * 30000.0), -32768, 32767).astype(np.int16)
           for i, n in enumerate(lens)]
    ref = [ex.calc_hashprint(p.astype(np.float32) / np.float32(32768.0)) for p in pcm]
    for p, r in zip(pcm, ref):
        assert np.array_equal(ex.calc_hashprint_pcm16(p), r)
    # batch on the device: even offsets (8-byte aligned floats), as the float batch call requires
    padded = [np.concatenate([p, np.zeros(len(p) & 1, dtype=np.int16)]) for p in pcm]
    so = np.zeros(len(pcm) + 1, dtype=np.int64)
    np.cumsum([len(p) for p in padded], out=so[1:])
    d_pcm = torch.from_numpy(np.concatenate(padded)).cuda()
    # each track is taken with its padded length: compare against the float path on the same padded buffers
    refp = [ex.calc_hashprint(p.astype(np.float32) / np.float32(32768.0)) for p in padded]
    d_hp = torch.zeros(sum(len(r) for r in refp), dtype=torch.int64, device="cuda")
    ex.calc_hashprint_pcm16_batch_device(d_pcm.data_ptr(), so, d_hp.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = d_hp.cpu().numpy().view(np.uint64)
    assert np.array_equal(got, np.concatenate(refp))


@pytest.mark.parametrize("seconds,sr", [(6.0, 44100), (30.0, 22050), (20.0, 44100)])
def test_cqt_window_switch_vs_oracle(ctx, seconds, sr):
    """the band window convention ("window","hann" at /root/reference/include/hpfw/spectrum/cqt.h:58)
    is a switch, not a constant. Both settings must match the oracle's same-named switch to the usual tolerance, and they
    must differ from each other (otherwise the switch does nothing)."""
    audio = synth.synth_track(int(seconds) + sr + 5, seconds, sr)
    mags = {}
    try:
        for wid, name in ((1, "symmetric"), (0, "periodic")):
            check(ctx._lib.hpfw_set_cqt_window(ctx.handle, wid))
            ref = nsgcq.nsgcq_magnitude(audio, window=name)
            mag = _cqt(ctx, audio, magnitude=True)
            assert np.max(np.abs(mag - ref)) <= 1e-5 * np.abs(ref).max(), name
            ref_db = nsgcq.amplitude_to_db(ref)
            db = _cqt(ctx, audio)
            above = ref_db > -79.0
            assert np.max(np.abs(db[above] - ref_db[above])) <= 0.01, name
            mags[name] = mag
    finally:
        check(ctx._lib.hpfw_set_cqt_window(ctx.handle, 0))
    # the two conventions differ by O(1 / Lg) per tap: far above the 1e-5 comparison tolerance at these lengths
    assert np.max(np.abs(mags["symmetric"] - mags["periodic"])) > 1e-4 * np.abs(mags["periodic"]).max()


def test_cqt_window_rejects_unknown(ctx):
    from hpfw_b200 import HpfwError
    with pytest.raises(HpfwError):
        check(ctx._lib.hpfw_set_cqt_window(ctx.handle, 2))
