"""The C++ host API (include/hpfw/*.h, mirroring the reference's classes) and the re-exported par_collector_* C ABI.

CPU part: the headers compile as C++17 against the library and fail loudly without a GPU.
GPU part: BASELINE.json configs[0] end to end through the reference's own example flow — 10 synthetic 30 s tracks at
22.05 kHz, 10 pitch-shifted noisy 6 s queries, LiveSongIdentification::index()/search() — compared with the oracle chain
(oracle CQT -> reference-pinned hashprint restatement -> MemoryStorage::find restatement).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "hpfw_b200")


def _build_example(tmp, rate=None):
    exe = os.path.join(tmp, "live-id")
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "cpp", "live-id.cpp"), "-o", exe, "-L" + LIBDIR, "-lhpfw_b200",
           "-Wl,-rpath," + LIBDIR]
    if rate:
        cmd.insert(1, f"-DHPFW_EXAMPLE_RATE={rate}")
    subprocess.check_call(cmd)
    return exe


def _save_filters_cereal(path, filt_cm):
    """cache/filters.cereal layout (reference utils.h:77-90): int32 rows, int32 cols, column-major floats."""
    with open(path, "wb") as f:
        f.write(np.array([64, 2420], dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(filt_cm, dtype=np.float32).tobytes())     # [2420,64] C-order == 64x2420 col-major


def _write_wav(path, x, sr):
    x = np.ascontiguousarray(x, dtype=np.float32)
    hdr = b"RIFF" + np.uint32(36 + x.nbytes).tobytes() + b"WAVEfmt " + np.uint32(16).tobytes() + \
        np.uint16(3).tobytes() + np.uint16(1).tobytes() + np.uint32(sr).tobytes() + np.uint32(sr * 4).tobytes() + \
        np.uint16(4).tobytes() + np.uint16(32).tobytes() + b"data" + np.uint32(x.nbytes).tobytes()
    with open(path, "wb") as f:
        f.write(hdr)
        f.write(x.tobytes())


def test_headers_compile_and_fail_loudly_without_gpu(tmp_path):
    import torch
    exe = _build_example(str(tmp_path))
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    os.makedirs(tmp_path / "a")
    os.makedirs(tmp_path / "b")
    p = subprocess.run([exe, "a", "b"], cwd=tmp_path, capture_output=True, text=True)
    assert p.returncode == 1
    assert "no CPU fallback" in p.stderr


def test_cereal_compat_roundtrip_of_reference_layout(tmp_path, hashprint_golden):
    """A tiny C++ program reads filters.cereal / writes a DB dump with the product headers; numpy checks the bytes."""
    src = tmp_path / "t.cpp"
    src.write_text(r'''
#include <hpfw/io/cereal_compat.h>
#include <iostream>
int main(int argc, char** argv) {
    hpfw::Matrix<float> f;
    if (!hpfw::io::load_matrix(argv[1], f)) return 3;
    if (f.rows() != 64 || f.cols() != 2420) return 4;
    hpfw::io::save_matrix(argv[2], f);
    std::vector<hpfw::io::NamedHashprint> db{{"alpha", {1, 2, 3}}, {"", {}}, {"b", {0xFFFFFFFFFFFFFFFFull}}};
    hpfw::io::save_db(argv[3], db);
    auto back = hpfw::io::load_db(argv[3]);
    if (back != db) return 5;
    std::cout << f(3, 7) << std::endl;
    return 0;
}''')
    exe = tmp_path / "t"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    filt = hashprint_golden["filters"]
    _save_filters_cereal(tmp_path / "in.cereal", filt)
    out = subprocess.check_output([str(exe), str(tmp_path / "in.cereal"), str(tmp_path / "out.cereal"),
                                   str(tmp_path / "db.cereal")], text=True)
    assert open(tmp_path / "in.cereal", "rb").read() == open(tmp_path / "out.cereal", "rb").read()
    assert abs(float(out) - float(filt[7 * 1 + 0, 3] if False else filt[7, 3])) < 1e-6     # F(3,7) = memory[7*64+3]
    raw = open(tmp_path / "db.cereal", "rb").read()
    exp = np.uint64(3).tobytes() + np.uint64(5).tobytes() + b"alpha" + np.uint64(3).tobytes() + \
        np.array([1, 2, 3], dtype=np.uint64).tobytes() + np.uint64(0).tobytes() + np.uint64(0).tobytes() + \
        np.uint64(1).tobytes() + b"b" + np.uint64(1).tobytes() + np.uint64(2 ** 64 - 1).tobytes()
    assert raw == exp


def test_wav_reader_formats_cpu(tmp_path):
    """include/hpfw/io/wav.h on every sample format it accepts: the bulk paths (mono float32, mono 16-bit PCM, raw int16
    hand-over) and the general per-sample path (8/24/32-bit PCM, float64, stereo down-mix) decode the same audio to the
    same floats. Host-only."""
    src = tmp_path / "w.cpp"
    src.write_text(r'''
#include <hpfw/io/wav.h>
#include <cstdio>
int main(int argc, char** argv) {
    for (int i = 1; i < argc; ++i) {
        auto w = hpfw::io::read_wav(argv[i]);
        auto k = hpfw::io::read_wav(argv[i], true);
        double s = 0, s2 = 0;
        for (float v : w.mono) { s += v; s2 += double(v) * v; }
        long long p = 0;
        for (auto v : k.pcm16) p += v;
        std::printf("%d %zu %.9g %.9g %zu %zu %lld\n", w.sample_rate, w.mono.size(), s, s2, k.mono.size(), k.pcm16.size(), p);
    }
    return 0;
}''')
    exe = tmp_path / "w"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    rng = np.random.default_rng(0)
    pcm = rng.integers(-32768, 32768, size=1001).astype(np.int16)
    f32 = pcm.astype(np.float32) / np.float32(32768.0)

    def wav(path, fmt, bits, payload, channels=1, sr=44100):
        hdr = b"RIFF" + np.uint32(36 + len(payload)).tobytes() + b"WAVEfmt " + np.uint32(16).tobytes() + \
            np.uint16(fmt).tobytes() + np.uint16(channels).tobytes() + np.uint32(sr).tobytes() + \
            np.uint32(sr * channels * bits // 8).tobytes() + np.uint16(channels * bits // 8).tobytes() + \
            np.uint16(bits).tobytes() + b"data" + np.uint32(len(payload)).tobytes()
        open(path, "wb").write(hdr + payload)
        return str(path)

    p24 = b"".join(int(v << 8).to_bytes(3, "little", signed=True) for v in pcm.tolist())
    files = [
        wav(tmp_path / "p16.wav", 1, 16, pcm.tobytes()),
        wav(tmp_path / "f32.wav", 3, 32, f32.tobytes()),
        wav(tmp_path / "f64.wav", 3, 64, f32.astype(np.float64).tobytes()),
        wav(tmp_path / "p24.wav", 1, 24, p24),
        wav(tmp_path / "p32.wav", 1, 32, (pcm.astype(np.int32) << 16).tobytes()),
        wav(tmp_path / "p16s.wav", 1, 16, np.stack([pcm, pcm], axis=1).tobytes(), channels=2),
    ]
    rows = [ln.split() for ln in subprocess.check_output([str(exe)] + files, text=True).strip().splitlines()]
    exp_s, exp_s2 = float(f32.astype(np.float64).sum()), float((f32.astype(np.float64) ** 2).sum())
    for r in rows:
        assert int(r[0]) == 44100 and int(r[1]) == len(pcm)
        assert abs(float(r[2]) - exp_s) <= 1e-6 * max(1.0, abs(exp_s)) and abs(float(r[3]) - exp_s2) <= 1e-6 * exp_s2
    # keep_pcm16: only the mono 16-bit file is handed over raw (and then not converted on the host)
    assert [int(r[5]) for r in rows] == [len(pcm), 0, 0, 0, 0, 0]
    assert int(rows[0][4]) == 0 and int(rows[0][6]) == int(pcm.astype(np.int64).sum())
    assert all(int(r[4]) == len(pcm) for r in rows[1:])


@pytest.mark.gpu
def test_live_id_example_config0(tmp_path, hashprint_golden):
    import oracle
    from oracle import nsgcq
    from hpfw_b200 import synth
    sr = 22050
    exe = _build_example(str(tmp_path), rate=sr)
    os.makedirs(tmp_path / "original")
    os.makedirs(tmp_path / "slices")
    filt = hashprint_golden["filters"]
    _save_filters_cereal(tmp_path / "filters.cereal", filt)
    tracks, names = [], []
    for i in range(10):
        t = synth.synth_track(3000 + i, 30.0, sr)
        tracks.append(t)
        names.append(f"track{i:02d}")
        _write_wav(tmp_path / "original" / f"track{i:02d}.wav", t, sr)
    queries = []
    for i in range(10):
        q, start = synth.synth_query(tracks[i], 4000 + i, 6.0, sr)
        queries.append(q)
        _write_wav(tmp_path / "slices" / f"q{i:02d}_track{i:02d}.wav", q, sr)
    p = subprocess.run([exe, "original", "slices", "filters.cereal"], cwd=tmp_path, capture_output=True, text=True,
                       timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln[3:].split() for ln in p.stdout.splitlines() if ln.startswith("=> ") and not ln.startswith("=> Finding")]
    results, summary = lines[:-1], lines[-1]
    assert len(results) == 10
    # index() re-learns the filters from the indexed tracks, as the reference's prepare() does (parallel_collector.h:111):
    # the oracle chain below uses the filters the run left in cache/filters.cereal
    hdr = np.fromfile(tmp_path / "cache" / "filters.cereal", dtype=np.int32, count=2)
    assert tuple(hdr) == (64, 2420)
    learned = np.fromfile(tmp_path / "cache" / "filters.cereal", dtype=np.float32, offset=8).reshape(2420, 64)
    assert not np.array_equal(learned, filt)
    filt = learned
    # oracle chain on the same audio
    hps = [oracle.hashprint_from_spectrogram(nsgcq.spectrogram(t), filt) for t in tracks]
    words, offs = oracle.pack_db(hps)
    for i, q in enumerate(queries):
        hq = oracle.hashprint_from_spectrogram(nsgcq.spectrogram(q), filt)
        tr, cnt, off = oracle.find(words, offs, hq)
        name, gcnt, goff = results[i][0], int(results[i][1]), int(results[i][2])
        assert name == names[tr] == names[i]                 # identical top-1 (and the right track)
        assert goff == off
        assert abs(gcnt - cnt) <= 2e-3 * 64 * len(hq)        # <= 0.1 % bits may differ in the query and in the track
    assert int(summary[0]) == 0 and float(summary[1]) == 1.0
    # cache layout is the reference's: spectros/<stem>, filters.cereal
    assert sorted(os.listdir(tmp_path / "cache" / "spectros")) == names
    spec = np.fromfile(tmp_path / "cache" / "spectros" / "track00", dtype=np.float32, offset=8).reshape(-1, 121)
    hdr = np.fromfile(tmp_path / "cache" / "spectros" / "track00", dtype=np.int32, count=2)
    assert tuple(hdr) == (121, 1210)
    ref = nsgcq.spectrogram(tracks[0])
    above = ref > -79
    assert np.max(np.abs(spec[above] - ref[above])) <= 0.01


@pytest.mark.gpu
def test_pyhpfw_wrapper_matches_cpp_path(tmp_path, hashprint_golden):
    """The reference's ctypes class against the re-exported par_collector_* symbols."""
    from hpfw_b200.pyhpfw import ParallelCollector
    from hpfw_b200 import HpfwError, synth
    sr = 44100
    cache = str(tmp_path / "cache") + "/"
    os.makedirs(cache + "spectros")
    _save_filters_cereal(cache + "filters.cereal", hashprint_golden["filters"])
    a = synth.synth_track(77, 6.0, sr)
    _write_wav(tmp_path / "a.wav", a, sr)
    pc = ParallelCollector()
    with pytest.raises(HpfwError):
        pc.calc_hashprint(str(tmp_path / "a.wav"))          # no filters yet: loud failure, not a crash
    pc.load(cache)
    hp = pc.calc_hashprint(str(tmp_path / "a.wav"))
    assert hp.dtype == np.uint64 and len(hp) == 385
    import oracle
    from oracle import nsgcq
    ref = oracle.hashprint_from_spectrogram(nsgcq.spectrogram(a), hashprint_golden["filters"])
    diff = int(np.unpackbits((hp ^ ref).view(np.uint8)).sum())
    assert diff <= 1e-3 * 64 * len(ref)
    # prepare() re-learns the filters from the indexed files (parallel_collector.h:111) and saves them
    got = pc.prepare([str(tmp_path / "a.wav"), str(tmp_path / "missing.wav")])     # unreadable file is skipped
    assert len(got) == 1 and got[0][0] == "a" and len(got[0][1]) == 385
    assert np.array_equal(pc.calc_hashprint(str(tmp_path / "a.wav")), got[0][1])
    learned = np.fromfile(cache + "filters.cereal", dtype=np.float32, offset=8).reshape(2420, 64)
    ref2 = oracle.hashprint_from_spectrogram(nsgcq.spectrogram(a), learned)
    diff2 = int(np.unpackbits((got[0][1] ^ ref2).view(np.uint8)).sum())
    assert diff2 <= 1e-3 * 64 * len(ref2)


def _write_wav16(path, pcm, sr, channels=1):
    x = np.ascontiguousarray(pcm, dtype=np.int16)
    hdr = b"RIFF" + np.uint32(36 + x.nbytes).tobytes() + b"WAVEfmt " + np.uint32(16).tobytes() + \
        np.uint16(1).tobytes() + np.uint16(channels).tobytes() + np.uint32(sr).tobytes() + \
        np.uint32(sr * 2 * channels).tobytes() + np.uint16(2 * channels).tobytes() + np.uint16(16).tobytes() + \
        b"data" + np.uint32(x.nbytes).tobytes()
    with open(path, "wb") as f:
        f.write(hdr)
        f.write(x.tobytes())


@pytest.mark.gpu
def test_pcm16_wav_takes_the_device_conversion_and_matches_float(tmp_path, hashprint_golden):
    """A mono 16-bit PCM WAV is shipped to the GPU as it is (hpfw_cqt_spectrogram_pcm16, sample / 32768 on the device); the
    same samples as a float32 WAV, and as a stereo file with two equal channels (host down-mix path), give the same
    hashprint bit for bit."""
    from hpfw_b200.pyhpfw import ParallelCollector
    from hpfw_b200 import synth
    sr = 44100
    cache = str(tmp_path / "cache") + "/"
    os.makedirs(cache + "spectros")
    _save_filters_cereal(cache + "filters.cereal", hashprint_golden["filters"])
    pcm = np.clip(np.round(synth.synth_track(78, 6.0, sr) * 30000.0), -32768, 32767).astype(np.int16)
    _write_wav16(tmp_path / "a16.wav", pcm, sr)
    _write_wav(tmp_path / "af.wav", pcm.astype(np.float32) / np.float32(32768.0), sr)
    _write_wav16(tmp_path / "a16s.wav", np.stack([pcm, pcm], axis=1), sr, channels=2)
    pc = ParallelCollector()
    pc.load(cache)
    h16 = pc.calc_hashprint(str(tmp_path / "a16.wav"))
    hf = pc.calc_hashprint(str(tmp_path / "af.wav"))
    hs = pc.calc_hashprint(str(tmp_path / "a16s.wav"))
    assert len(h16) == 385
    assert np.array_equal(h16, hf) and np.array_equal(h16, hs)
