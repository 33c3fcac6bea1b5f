"""The device-resident index()/search() pipeline behind the reference's C++ API (hpfw_xs_*, xstream.cu; ParallelCollector in
include/hpfw/core/parallel_collector.h) against the single-file calls of the same library and against the oracle chain.

Reference flow being replaced: ParallelCollector::prepare = preprocess (fan-out over files, covariance accumulate, cache
file per spectrogram) + collect_fingerprints (every cached spectrogram re-read and hashed),
/root/reference/include/hpfw/core/parallel_collector.h:48-52, 82-137; LiveSongIdentification::search, live_song_id.h:35-54.
"""
import json
import os
import subprocess

import numpy as np
import pytest

from hpfw_b200.bench_data import write_wav_f32 as _write_wav, write_wav_pcm16 as _write_wav16

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "hpfw_b200")

pytestmark = pytest.mark.gpu

SR = 44100


def _pcm(seed, seconds):
    from hpfw_b200 import synth
    return np.clip(np.round(synth.synth_track(seed, seconds, SR) * 30000.0), -32768, 32767).astype(np.int16)


def _make_library(tmp_path):
    """7 files: mixed lengths and sample formats, one too short for a hashprint word, one unreadable."""
    files, pcm = [], {}
    for i, (secs, kind) in enumerate([(6.0, "p16"), (9.5, "f32"), (6.0, "p16"), (20.0, "p16"), (7.25, "f32")]):
        x = _pcm(500 + i, secs)
        name = f"t{i}_{kind}"
        path = str(tmp_path / f"{name}.wav")
        if kind == "p16":
            _write_wav16(path, x, SR)
        else:
            _write_wav(path, x.astype(np.float32) / np.float32(32768.0), SR)
        files.append(path)
        pcm[name] = x
    _write_wav16(str(tmp_path / "short.wav"), _pcm(600, 0.8), SR)       # < 100 spectrogram columns: skipped, logged
    files.insert(2, str(tmp_path / "short.wav"))
    files.append(str(tmp_path / "does_not_exist.wav"))
    return files, pcm


def test_prepare_batched_equals_single_file_calls(tmp_path, hashprint_golden, capfd):
    """prepare() through the pipeline (decode threads -> pinned ring -> lanes -> resident spectrograms -> batched projection)
    returns, for every file, exactly the words calc_hashprint() gives for that file alone with the same filters; bad files
    are skipped like the reference does; names are the stems, sorted; cache files hold the spectrograms."""
    from hpfw_b200.pyhpfw import ParallelCollector
    import oracle
    from oracle import nsgcq
    files, pcm = _make_library(tmp_path)
    cache = str(tmp_path / "cache") + "/"
    pc = ParallelCollector()
    pc.load(cache)
    got = pc.prepare(files)
    err = capfd.readouterr().err
    assert "short.wav" in err and "does_not_exist.wav" in err
    assert [g[0] for g in got] == sorted(pcm)
    for name, hp in got:
        single = pc.calc_hashprint(str(tmp_path / f"{name}.wav"))
        assert np.array_equal(hp, single), name
    # against the oracle chain with the filters this run learned
    learned = np.fromfile(cache + "filters.cereal", dtype=np.float32, offset=8).reshape(2420, 64)
    for name, hp in got:
        ref = oracle.hashprint_from_spectrogram(nsgcq.spectrogram(pcm[name].astype(np.float32) / 32768.0), learned)
        diff = int(np.unpackbits((hp ^ ref).view(np.uint8)).sum())
        assert len(hp) == len(ref) and diff <= 1e-3 * 64 * len(ref), (name, diff)
    # the cache directory is the reference's (background writers have been flushed by save / the next call)
    pc.save(cache)
    del pc
    assert sorted(os.listdir(cache + "spectros")) == sorted(pcm)
    spec = np.fromfile(cache + "spectros/t3_p16", dtype=np.float32, offset=8).reshape(-1, 121)
    ref = nsgcq.spectrogram(pcm["t3_p16"].astype(np.float32) / 32768.0)
    above = ref > -79
    assert spec.shape == ref.shape and np.max(np.abs(spec[above] - ref[above])) <= 0.01


def test_prepare_rehashes_older_cache_files_and_survives_save_load(tmp_path, hashprint_golden):
    """collect_fingerprints hashes EVERY spectrogram in cache/spectros (parallel_collector.h:119-134), not only this run's
    files; and the ADVICE r1 case prepare() -> save(other dir) -> load(other dir) -> calc_hashprint keeps the learned state."""
    from hpfw_b200.pyhpfw import ParallelCollector
    files, pcm = _make_library(tmp_path)
    cache = str(tmp_path / "cache") + "/"
    pc = ParallelCollector()
    pc.load(cache)
    first = dict(pc.prepare(files[:2]))                    # t0, t1
    assert sorted(first) == ["t0_p16", "t1_f32"]
    second = dict(pc.prepare([files[4]]))                   # t3 now; t0 and t1 come back from their cache files
    assert sorted(second) == ["t0_p16", "t1_f32", "t3_p16"]
    for name in second:                                     # all hashed with the filters of the second run
        assert np.array_equal(second[name], pc.calc_hashprint(str(tmp_path / f"{name}.wav"))), name
    other = str(tmp_path / "elsewhere") + "/"
    pc.save(other)
    assert os.path.exists(other + "filters.cereal") and os.path.exists(other + "accum_cov.cereal")
    pc.load(other)
    assert np.array_equal(pc.calc_hashprint(files[0]), second["t0_p16"])
    pc2 = ParallelCollector(other)                          # a fresh collector picks the saved state up
    assert np.array_equal(pc2.calc_hashprint(files[0]), second["t0_p16"])


def test_arena_spill_goes_through_the_cache_files(tmp_path, monkeypatch):
    """With an HBM budget smaller than the library's spectrograms, prepare() spills through cache/spectros (the reference's
    normal flow) and returns the same hashprints as the fully resident run."""
    from hpfw_b200.pyhpfw import ParallelCollector
    files, pcm = _make_library(tmp_path)
    base = ParallelCollector()
    base.load(str(tmp_path / "c0") + "/")
    want = dict(base.prepare(files))
    del base
    # spectrograms: 6 s = 484 columns x 121 x 4 B = 0.23 MB, 9.5 s = 0.37 MB, 20 s = 0.78 MB; the library needs 1.9 MB. A budget of
    # 1.25 MB in 0.5 MB chunks forces two spills
    monkeypatch.setenv("HPFW_XS_ARENA_BYTES", str(5 << 18))
    monkeypatch.setenv("HPFW_XS_CHUNK_BYTES", str(1 << 19))
    pc = ParallelCollector()
    pc.load(str(tmp_path / "c1") + "/")
    got = dict(pc.prepare(files))
    assert sorted(got) == sorted(want)
    for name in want:
        assert np.array_equal(got[name], want[name]), name


def _build_bench(tmp):
    exe = os.path.join(tmp, "bench-liveid")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "cpp", "bench-liveid.cpp"), "-o", exe, "-L" + LIBDIR,
                           "-lhpfw_b200", "-lpthread", "-Wl,-rpath," + LIBDIR])
    return exe


def test_cpp_index_and_search_device_path(tmp_path, hashprint_golden):
    """LiveSongIdentification::index()/search() through the compiled C++ bench binary: index() builds the DB
    device-to-device; search() extracts all query files in one batch and matches them in HBM; every query finds its track."""
    from hpfw_b200 import synth
    from hpfw_b200.bench_data import write_db_dump
    exe = _build_bench(str(tmp_path))
    os.makedirs(tmp_path / "tracks")
    os.makedirs(tmp_path / "queries")
    os.makedirs(tmp_path / "cache" / "spectros")
    audio = []
    for i in range(6):
        a = synth.synth_track(800 + i, 30.0, SR)
        audio.append(a)
        _write_wav16(str(tmp_path / "tracks" / f"song{i:02d}.wav"), np.round(a * 30000.0).astype(np.int16), SR)
    p = subprocess.run([exe, "index", "tracks", "2"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    rec = json.loads(p.stdout.strip().splitlines()[-1])
    assert rec["leg"] == "index" and rec["files"] == 6 and rec["db_tracks"] == 6 and rec["frames_per_s"] > 0
    # search leg: the DB comes from a dump written here with the filters index() learned
    import hpfw_b200
    learned = np.fromfile(tmp_path / "cache" / "filters.cereal", dtype=np.float32, offset=8).reshape(2420, 64)
    ctx = hpfw_b200.Context(0)
    ex = hpfw_b200.HashprintExtractor(ctx)
    ex.set_filters(learned)
    names = [f"song{i:02d}" for i in range(6)]
    hps = [ex.calc_hashprint_pcm16(np.round(a * 30000.0).astype(np.int16)) for a in audio]
    write_db_dump(str(tmp_path / "db.cereal"), names, hps)
    expect = []
    for q in range(12):
        qa, _ = synth.synth_query(audio[q % 6], 900 + q, 6.0, SR, max_semitones=0.25)
        fn = f"q{q:02d}_{names[q % 6]}.wav"
        _write_wav16(str(tmp_path / "queries" / fn), np.clip(np.round(qa * 30000.0), -32768, 32767).astype(np.int16), SR)
        expect.append(f"{fn} {names[q % 6]}")
    _write_wav16(str(tmp_path / "queries" / "zz_too_short.wav"), _pcm(1, 0.5), SR)
    (tmp_path / "expect.txt").write_text("\n".join(expect) + "\n")
    p = subprocess.run([exe, "search", "db.cereal", "queries", "2", "expect.txt"], cwd=tmp_path, capture_output=True,
                       text=True, timeout=300)
    rec = json.loads(p.stdout.strip().splitlines()[-1])
    # the too-short file fails alone (per-query error), every real query finds its track
    assert rec["queries"] == 13 and rec["failed"] == 1 and rec["wrong"] == 0 and rec["checked"] == 12, (rec, p.stderr[-800:])
    assert "zz_too_short.wav" in p.stderr
    ctx.close()
